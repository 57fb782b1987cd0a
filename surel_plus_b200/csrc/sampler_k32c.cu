// Sampler instantiations with 32-bit (node << OB | order) keys, part c of the keys-per-lane list.
#define SUBG_KEY_T uint32_t
#define SUBG_LAUNCH_NAME launch_gset_sample_k32c
#define SUBG_EPL_CASES CASE(25) CASE(29) CASE(33)
#include "sampler_launch.inc"
