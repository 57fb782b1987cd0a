// gset_hash_kernel instantiations with 64-bit entries / sort keys (graphs with more nodes than a 32-bit key can hold).
#define SUBG_HASH_KEY_T unsigned long long
#define SUBG_HASH_LAUNCH_NAME launch_gset_hash_k64
#include "sampler_hash_launch.inc"
