// Sampler instantiations with 64-bit (node << OB | order) keys, part a of the keys-per-lane list.
#define SUBG_KEY_T uint64_t
#define SUBG_LAUNCH_NAME launch_gset_sample_k64a
#define SUBG_EPL_CASES CASE(3) CASE(4) CASE(5) CASE(7) CASE(8) CASE(9) CASE(11) CASE(13)
#include "sampler_launch.inc"
