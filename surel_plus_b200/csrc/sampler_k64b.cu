// Sampler instantiations with 64-bit (node << OB | order) keys, part b of the keys-per-lane list.
#define SUBG_KEY_T uint64_t
#define SUBG_LAUNCH_NAME launch_gset_sample_k64b
#define SUBG_EPL_CASES CASE(15) CASE(17) CASE(19) CASE(21)
#include "sampler_launch.inc"
