// Sampler instantiations with 32-bit (node << OB | order) keys: graphs with N < 2^(32-OB).
#define SUBG_KEY_T uint32_t
#define SUBG_LAUNCH_NAME launch_gset_sample_k32
#include "sampler_launch.inc"
