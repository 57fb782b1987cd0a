// gset_hash_kernel instantiations with 32-bit entries / sort keys (node bits + order bits <= 32).
#define SUBG_HASH_KEY_T uint32_t
#define SUBG_HASH_LAUNCH_NAME launch_gset_hash_k32
#include "sampler_hash_launch.inc"
