// Sampler instantiations with 32-bit (node << OB | order) keys, part d of the keys-per-lane list.
#define SUBG_KEY_T uint32_t
#define SUBG_LAUNCH_NAME launch_gset_sample_k32d
#define SUBG_EPL_CASES CASE(41) CASE(49) CASE(63)
#include "sampler_launch.inc"
