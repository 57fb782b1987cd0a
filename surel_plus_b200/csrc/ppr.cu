// PPR set sampler (ACL forward push + top-k) -- sampler/pprgo.py:9-111.
#include "common.cuh"

namespace subg {

int ppr_topk_impl(const Graph *g, const int32_t *seeds_hd, int64_t n, float alpha, float eps, int topk,
                  int normalization, int encoder, cudaStream_t st, SpG **out) {
    (void)g; (void)seeds_hd; (void)n; (void)alpha; (void)eps; (void)topk; (void)normalization; (void)encoder; (void)st; (void)out;
    return fail(SUBG_ERR_UNSUPPORTED, "subg_ppr_topk: not built yet");
}

}  // namespace subg
