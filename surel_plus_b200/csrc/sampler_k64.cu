// Sampler instantiations with 64-bit keys: node ids that do not fit beside the order field.
#define SUBG_KEY_T unsigned long long
#define SUBG_LAUNCH_NAME launch_gset_sample_k64
#include "sampler_launch.inc"
