// Launcher of the second-generation large-M int4 kernel (gemm_w4p.cuh): CTA pairs, weight operand in tensor memory.
#include <cstring>

#include "api_qbits.cuh"
#include "gemm_w4p.cuh"

namespace qb {

template <class Cfg>
static int launch_w4p_cfg(const CUtensorMap& tw, const CUtensorMap& tx, const StoreMaps& sm, const W4PParams& p,
                          uint32_t idesc, cudaStream_t stream) {
  int rc = ensure_dyn_smem<gemm_w4p_kernel<Cfg>>(Cfg::SMEM_BYTES);
  if (rc != OK) return rc;
  const int tiles = p.num_tok_blocks * p.num_feat_blocks;
  const int pairs = current_sm_count() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  gemm_w4p_kernel<Cfg><<<grid, Cfg::NTHREADS, Cfg::SMEM_BYTES, stream>>>(tw, tx, sm, p, idesc);
  return check_cuda(cudaGetLastError(), "gemm_w4p_kernel launch");
}

int launch_w4p(const QbitsArgs& q) {
  W4PParams p{};
  p.scale = q.scale;
  p.shift = q.shift;
  p.bias = q.bias;
  p.g = q.g;
  p.ld = static_cast<int>(q.ld);
  p.col0 = static_cast<int>(q.col0);
  p.M = static_cast<int>(q.m);
  p.N = static_cast<int>(q.n);
  p.K = static_cast<int>(q.k);
  p.group = q.group;
  p.group_log2 = q.group_log2;
  p.num_tok_blocks = static_cast<int>((q.m + 255) / 256);
  p.num_feat_blocks = static_cast<int>((q.n / 2 + 127) / 128);
  p.trace = debug_trace();
  CUtensorMap tw, tx;
  StoreMaps sm;
  std::memset(&sm, 0, sizeof(sm));
  int rc = make_tmap_2d(&tw, q.packed, DT_U8, q.n / 2, q.k, 64);
  if (rc != OK) return rc;
  rc = make_tmap_2d(&tx, q.a, q.dtype, q.m, q.k, 128);
  if (rc != OK) return rc;
  for (int i = 0; i < q.g.n_out; ++i) {
    rc = make_tmap_2d_view(&sm.m[i], q.g.out_peer[i], q.dtype, q.m, q.ld, q.ld, 64, 32, false);
    if (rc != OK) return rc;
  }
  set_kernel_family(1);
  const uint32_t fmt = (q.dtype == DT_BF16) ? 1u : 0u;
  const uint32_t idesc = umma_idesc(1u, fmt, fmt, 256u, 256u);
  const bool zp = q.shift_is_int != 0;
  if (q.dtype == DT_BF16) {
    if (zp) return launch_w4p_cfg<W4PCfg<__nv_bfloat16, true>>(tw, tx, sm, p, idesc, q.stream);
    return launch_w4p_cfg<W4PCfg<__nv_bfloat16, false>>(tw, tx, sm, p, idesc, q.stream);
  }
  if (zp) return launch_w4p_cfg<W4PCfg<__half, true>>(tw, tx, sm, p, idesc, q.stream);
  return launch_w4p_cfg<W4PCfg<__half, false>>(tw, tx, sm, p, idesc, q.stream);
}

}  // namespace qb
