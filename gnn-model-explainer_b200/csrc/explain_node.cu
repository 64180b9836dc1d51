// explain_node.cu -- K2: the persistent per-node mask-optimisation kernel (node mode), v2.
//
// One CTA owns one explained node for ALL epochs: mask build A (.) sym(sigmoid(M)), the reference's
// 3-layer GCN forward ((A_m H) W + b -> row L2-normalise -> ReLU), softmax / -log p[gt], the
// size / entropy / Laplacian / feature-size regularisers, the hand-derived backward to dL/dM and
// dL/dF, and the Adam step, with every array resident in shared memory (or, for tasks that do not
// fit 227 KB, in a per-CTA global-memory slab that stays in L2).  Replaces, for the default
// hyper-parameters, explainer/explain.py:137-146 (epoch loop) + :665-715 (ExplainModule.forward)
// + :740-808 (loss) + autograd + torch.optim.Adam, and models.py:58-80,230-267,363-376.
//
// What makes it cheaper than the dense reference (exact, not approximate):
//   * M, m, v live only on the E_d directed edges of the k-hop sub-adjacency: every term of
//     dL/dM_ij is local to (i,j)/(j,i), off-edge entries never reach the returned mask.
//   * Only ONE row of logits carries loss, so layer l is needed only for nodes within L-l hops
//     of the explained node; with nodes relabelled in (distance, degree) order every layer's row
//     set is a prefix [0,n_{L-l}) and the backward touches the same prefixes.
//   * dL/dF needs sum_i dZ1[i] (.) U[i] with U = A_m X kept from the forward, so the layer-1
//     transpose aggregation disappears.
//   * The returned mask is the one built in the forward of the LAST epoch (explain.py:694,209),
//     i.e. after num_epochs-1 updates; the last backward/Adam step is unobservable and skipped.
//
// Work mapping (v2, driven by the ncu profile of v1 in profiles/r01a_v1_kernel_summary.md):
//   * sparse aggregations: a warp is cut into groups of W/4 lanes; each group owns ONE row and
//     walks its edges with float4 shared-memory loads (rows of similar degree are adjacent thanks
//     to the plan's ordering); rows longer than kLongRow edges are split across the whole warp.
//   * dense 20x20 / dx20 products, row normalisation and their backward: ONE THREAD PER ROW,
//     accumulators in registers, weight rows broadcast from shared memory as float4.
//   * edge phase: one thread per undirected edge (both directions), sigmoid cached between epochs.
// Phases per epoch (one __syncthreads each): F1 | F2 | S (row r: layer 3 + readout + softmax +
// layer-3 backward, one warp) | B2 | B1 | P.
#include "gnnx_internal.cuh"

namespace {

constexpr int kLongRow = 32;  // rows with more edges than this are aggregated by a whole warp

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// Philox4x32-10 (Salmon et al. 2011), used only for GX_INIT_PHILOX.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t node, uint32_t slot) {
  uint32_t r[4];
  philox4x32_10(slot, node, 0x67u, 0x6e78u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const float u1 = ((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2);
}

template <typename IdxT> struct IdxTraits;
template <> struct IdxTraits<uint16_t> { static constexpr uint16_t kNone = 0xFFFFu; };
template <> struct IdxTraits<int32_t> { static constexpr int32_t kNone = -1; };

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 relu4(float4 v) {
  return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
}
__device__ __forceinline__ void fma4(float4& acc, float s, const float4 v) {
  acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}

// dot of two length-(4*n4) vectors / dot(a, relu(b))
__device__ __forceinline__ float dot_v4(const float* a, const float* b, int n4) {
  float s = 0.f;
  for (int k = 0; k < n4; ++k) {
    const float4 x = ld4(a + 4 * k), y = ld4(b + 4 * k);
    s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
  }
  return s;
}
__device__ __forceinline__ float dot_relu_v4(const float* a, const float* b, int n4) {
  float s = 0.f;
  for (int k = 0; k < n4; ++k) {
    const float4 x = ld4(a + 4 * k), y = relu4(ld4(b + 4 * k));
    s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
  }
  return s;
}

struct ExplainArgs {
  const int32_t* order;
  int32_t ntasks;
  int32_t* counter;
  float* gws;
  int64_t gws_stride_words;
  GxGraphDev g;
  GxModelDev m;
  GxHparamsDev hp;
  GxPlanArrays plan;
  const float* m0;
  float* out_mask;
  float* out_feat;
};

// ---------------------------------------------------------------------------------------------
// Sparse aggregation  dst[i] = sum_{e in row i (, col < col_limit)} a[e] * f(src[col[e]])  for the rows
// [row_b, row_e) of a warp's block.  W4 = row width in float4 (<= 32), lanes are cut into groups of W4.
//   kRelu   : f = relu
//   mode    : 0 = only rows with <= kLongRow edges (row per lane group)
//             1 = only rows with  > kLongRow edges, rows dealt across warps (edges across groups)
// ---------------------------------------------------------------------------------------------
template <typename IdxT, bool kRelu>
__device__ __forceinline__ void gather_short_rows(int row_b, int row_e, int W4, int lane,
                                                  const IdxT* __restrict__ irp, const IdxT* __restrict__ icol,
                                                  const float* a, const float* src, int src_stride,
                                                  float* dst, int dst_stride, int col_limit) {
  const int grp = lane / W4, q = lane - grp * W4;
  const int epi = 32 / W4;
  if (grp >= epi) return;
  for (int i = row_b + grp; i < row_e; i += epi) {
    const int r0 = irp[i], r1 = irp[i + 1];
    if (r1 - r0 > kLongRow) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = r0; e < r1; ++e) {
      const int c = icol[e];
      if (c >= col_limit) break;
      float4 v = ld4(src + c * src_stride + 4 * q);
      if (kRelu) v = relu4(v);
      fma4(acc, a[e], v);
    }
    st4(dst + i * dst_stride + 4 * q, acc);
  }
}

template <typename IdxT, bool kRelu>
__device__ __forceinline__ void gather_long_row(int i, int W4, int lane, const IdxT* __restrict__ irp,
                                                const IdxT* __restrict__ icol, const float* a,
                                                const float* src, int src_stride, float* dst, int dst_stride,
                                                int col_limit, float* scratch /* >= 128 floats per warp */) {
  const int grp = lane / W4, q = lane - grp * W4;
  const int epi = 32 / W4;
  const int r0 = irp[i], r1 = irp[i + 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (grp < epi) {
    for (int e = r0 + grp; e < r1; e += epi) {
      const int c = icol[e];
      if (c >= col_limit) break;  // columns are partitioned by level: once past the limit, all later ones are too
      float4 v = ld4(src + c * src_stride + 4 * q);
      if (kRelu) v = relu4(v);
      fma4(acc, a[e], v);
    }
    st4(scratch + lane * 4, acc);
  }
  __syncwarp();
  if (grp == 0) {
    float4 t = acc;
    for (int g2 = 1; g2 < epi; ++g2) {
      const float4 o = ld4(scratch + (g2 * W4 + q) * 4);
      t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
    }
    st4(dst + i * dst_stride + 4 * q, t);
  }
  __syncwarp();
}

template <bool kShared, typename IdxT, int HID, int EMB, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) explain_node_kernel(const ExplainArgs A) {
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ int s_task;
  __shared__ GxLayout sL;
  __shared__ int s_long[2];  // number of long rows among [0,n2) and among [0,n1)
  static_assert(HID % 4 == 0 && EMB % 4 == 0, "hidden widths must be multiples of 4");
  constexpr IdxT kNone = IdxTraits<IdxT>::kNone;
  constexpr int HS = HID;            // row stride of the hidden-width arrays
  constexpr int H4 = HID / 4;
  constexpr int PD = 2 * HID + EMB;  // pred_model input width (concat of the three layers)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  float* const base = kShared ? smem_dyn : (A.gws + (int64_t)blockIdx.x * A.gws_stride_words);
  const GxModelDev& m = A.m;
  const GxHparamsDev& hp = A.hp;
  const int d = m.d, C = m.C;

  for (;;) {
    if (tid == 0) s_task = atomicAdd(A.counter, 1);
    __syncthreads();
    const int qi = s_task;
    __syncthreads();
    if (qi >= A.ntasks) break;
    const int task_id = A.order[qi];
    const GxTask* __restrict__ Tp = A.plan.tasks + task_id;
    const int n = Tp->n, n1 = Tp->n1, n2 = Tp->n2, e1 = Tp->e1, np = Tp->npairs;
    const int gt = Tp->gt_label;
    const int64_t node_off = Tp->node_off, rp_off = Tp->rp_off, edge_off = Tp->edge_off, pair_off = Tp->pair_off;
    if (tid == 0) sL = gx_make_layout(n, n1, n2, e1, np, d, HID, EMB, C, nwarps, (int)sizeof(IdxT));
    __syncthreads();
    const int dp = sL.dp, D4 = dp / 4, TS = sL.ts;
    const int32_t* __restrict__ lo2gid = A.plan.lo2gid + node_off;
    const float nn = (float)n * (float)n;
    const float ent_over_nn = hp.c_ent / nn;
    const float lap_over_nn = hp.c_lap / nn;

    // ------------------------------------------------------------------ load
    {
      float* const X = base + sL.X; float* const W1s = base + sL.W1s; float* const W2s = base + sL.W2s; float* const W3s = base + sL.W3s;
      float* const bs = base + sL.bs; float* const sF = base + sL.sF; float* const Fm = base + sL.F; float* const mF = base + sL.mF;
      float* const vF = base + sL.vF; float* const gFp = base + sL.gFp; float* const a = base + sL.a; float* const lap2 = base + sL.lap2;
      float2* const MM = reinterpret_cast<float2*>(base + sL.M); float2* const mm = MM + np; float2* const vv = mm + np; float2* const SS = vv + np;
      IdxT* const icol = reinterpret_cast<IdxT*>(base + sL.icol); IdxT* const irp = reinterpret_cast<IdxT*>(base + sL.irp);
      IdxT* const pi = reinterpret_cast<IdxT*>(base + sL.pi); IdxT* const pj = reinterpret_cast<IdxT*>(base + sL.pj);
      IdxT* const ppij = reinterpret_cast<IdxT*>(base + sL.ppij); IdxT* const ppji = reinterpret_cast<IdxT*>(base + sL.ppji);
    for (int idx = tid; idx < n * dp; idx += nthreads) {
      const int i = idx / dp, f = idx - i * dp;
      X[idx] = f < d ? __ldg(A.g.feat + (int64_t)lo2gid[i] * d + f) : 0.f;
    }
    for (int idx = tid; idx < dp * HS; idx += nthreads) {
      const int f = idx / HS, c = idx - f * HS;
      W1s[idx] = f < d ? __ldg(m.W[0] + f * HID + c) : 0.f;
    }
    for (int idx = tid; idx < HID * HS; idx += nthreads) W2s[idx] = __ldg(m.W[1] + idx);
    for (int idx = tid; idx < HID * EMB; idx += nthreads) W3s[idx] = __ldg(m.W[2] + idx);
    for (int idx = tid; idx < HID; idx += nthreads) { bs[idx] = __ldg(m.b[0] + idx); bs[HID + idx] = __ldg(m.b[1] + idx); }
    for (int idx = tid; idx < EMB; idx += nthreads) bs[2 * HID + idx] = __ldg(m.b[2] + idx);
    for (int e = tid; e < e1; e += nthreads) icol[e] = (IdxT)A.plan.icol[edge_off + e];
    for (int i = tid; i <= n2; i += nthreads) irp[i] = (IdxT)A.plan.irowptr[rp_off + i];
    for (int f = tid; f < dp; f += nthreads) {
      sF[f] = 0.5f;  // sigmoid(0): feat_mask is initialised to 0 (explain.py:633-643)
      Fm[f] = 0.f; mF[f] = 0.f; vF[f] = 0.f;
    }
    for (int idx = tid; idx < nwarps * dp; idx += nthreads) gFp[idx] = 0.f;
    const float m0_std = sqrtf(2.0f / (float)n);  // gain('relu') * sqrt(2/(n+n)) (explain.py:647-651)
    for (int p = tid; p < np; p += nthreads) {
      const int i = A.plan.pair_i[pair_off + p], j = A.plan.pair_j[pair_off + p];
      const int pij = A.plan.pair_pij[pair_off + p], pji = A.plan.pair_pji[pair_off + p];
      const int oij = A.plan.pair_oij[pair_off + p], oji = A.plan.pair_oji[pair_off + p];
      pi[p] = (IdxT)i; pj[p] = (IdxT)j;
      ppij[p] = i < n2 ? (IdxT)pij : kNone;
      ppji[p] = j < n2 ? (IdxT)pji : kNone;
      float Mi, Mj;
      if (hp.init == GX_INIT_M0) {
        Mi = __ldg(A.m0 + edge_off + oij);
        Mj = __ldg(A.m0 + edge_off + oji);
      } else {
        Mi = 1.0f + m0_std * philox_normal(hp.seed, (uint32_t)Tp->node, (uint32_t)oij);
        Mj = 1.0f + m0_std * philox_normal(hp.seed, (uint32_t)Tp->node, (uint32_t)oji);
      }
      MM[p] = make_float2(Mi, Mj);
      mm[p] = make_float2(0.f, 0.f);
      vv[p] = make_float2(0.f, 0.f);
      const float Si = sigmoid_f(Mi), Sj = sigmoid_f(Mj);
      SS[p] = make_float2(Si, Sj);
      const float yi = (float)__ldg(A.g.pred_label + lo2gid[i]);
      const float yj = (float)__ldg(A.g.pred_label + lo2gid[j]);
      lap2[p] = lap_over_nn * (yi - yj) * (yi - yj);  // d/dA_ij + d/dA_ji of y^T (D - A) y / n^2
      const float a0 = 0.5f * (Si + Sj);  // explain.py:665-678
      if (i < n2) a[pij] = a0;
      if (j < n2) a[pji] = a0;
      if (hp.iters == 0) {
        A.out_mask[edge_off + oij] = a0;
        A.out_mask[edge_off + oji] = a0;
      }
    }
    }
    __syncthreads();
    // rows with more than kLongRow edges (hubs): explicit list (ascending row id), aggregated by whole
    // warps and dealt across warps; everything else goes through the lane-group path
    if (warp == 0) {
      const IdxT* const irp = reinterpret_cast<const IdxT*>(base + sL.irp);
      IdxT* const llist = reinterpret_cast<IdxT*>(base + sL.llist);
      int cnt = 0, cnt1 = 0;
      for (int b0 = 0; b0 < n2; b0 += 32) {
        const int i = b0 + lane;
        const bool lg = i < n2 && ((int)irp[i + 1] - (int)irp[i] > kLongRow);
        const uint32_t bal = __ballot_sync(0xffffffffu, lg);
        if (lg) llist[cnt + __popc(bal & ((1u << lane) - 1u))] = (IdxT)i;
        cnt += __popc(bal);
        cnt1 += __popc(__ballot_sync(0xffffffffu, lg && i < n1));
      }
      if (lane == 0) { s_long[0] = cnt; s_long[1] = cnt1; }
    }
    __syncthreads();
    const int nlongF1 = s_long[0];  // long rows among [0,n2)
    const int nlongF2 = s_long[1];  // long rows among [0,n1) (a prefix of the list)

    // ------------------------------------------------------------------ epochs
    for (int it = 1; it <= hp.iters; ++it) {
      // ---- F1: rows [0,n2): U = A_m X ; Y1 = (U . sF) W1 + b1 ; normalise           (models.py:70-78)
      {
      const IdxT* const irp = reinterpret_cast<const IdxT*>(base + sL.irp);
      const IdxT* const icol = reinterpret_cast<const IdxT*>(base + sL.icol);
      float* const a = base + sL.a;
      float* const X = base + sL.X;
      float* const U = base + sL.U;
      float* const zs = base + sL.zs + warp * 128;
      float* const bs = base + sL.bs;
      float* const sF = base + sL.sF;
      float* const W1s = base + sL.W1s;
      float* const Yh1 = base + sL.Yh1;
      float* const q1 = base + sL.q1;
      const IdxT* const llist = reinterpret_cast<const IdxT*>(base + sL.llist);
      if (nlongF1 > 0) {
        for (int k = warp; k < nlongF1; k += nwarps)
          gather_long_row<IdxT, false>((int)llist[k], D4, lane, irp, icol, a, X, dp, U, dp, n, zs);
        __syncthreads();
      }
      for (int rb = warp * 32; rb < n2; rb += nwarps * 32) {
        const int re = min(rb + 32, n2);
        gather_short_rows<IdxT, false>(rb, re, D4, lane, irp, icol, a, X, dp, U, dp, n);
        __syncwarp();
        const int i = rb + lane;
        if (i < re) {
          float acc[HID];
#pragma unroll
          for (int c4 = 0; c4 < H4; ++c4) {
            const float4 b = ld4(bs + 4 * c4);
            acc[4 * c4] = b.x; acc[4 * c4 + 1] = b.y; acc[4 * c4 + 2] = b.z; acc[4 * c4 + 3] = b.w;
          }
          for (int f4 = 0; f4 < D4; ++f4) {
            const float4 u = ld4(U + i * dp + 4 * f4);
            const float4 s = ld4(sF + 4 * f4);
            const float z[4] = {u.x * s.x, u.y * s.y, u.z * s.z, u.w * s.w};  // x * sigmoid(feat_mask) (explain.py:707)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float* wr = W1s + (4 * f4 + k) * HS;
#pragma unroll
              for (int c4 = 0; c4 < H4; ++c4) {
                const float4 w = ld4(wr + 4 * c4);
                acc[4 * c4] = fmaf(z[k], w.x, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(z[k], w.y, acc[4 * c4 + 1]);
                acc[4 * c4 + 2] = fmaf(z[k], w.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(z[k], w.w, acc[4 * c4 + 3]);
              }
            }
          }
          float ss = 0.f;
#pragma unroll
          for (int c = 0; c < HID; ++c) ss = fmaf(acc[c], acc[c], ss);
          const float q = fmaxf(sqrtf(ss), 1e-12f);  // F.normalize(p=2, dim=2), eps 1e-12
#pragma unroll
          for (int c4 = 0; c4 < H4; ++c4)
            st4(Yh1 + i * HS + 4 * c4, make_float4(acc[4 * c4] / q, acc[4 * c4 + 1] / q, acc[4 * c4 + 2] / q, acc[4 * c4 + 3] / q));
          q1[i] = q;
        }
      }
      }
      __syncthreads();
      // ---- F2: rows [0,n1): Y2 = (A_m relu(Yh1)) W2 + b2 ; normalise (aggregate lands in Yh2[i], then in place)
      {
      const IdxT* const irp = reinterpret_cast<const IdxT*>(base + sL.irp);
      const IdxT* const icol = reinterpret_cast<const IdxT*>(base + sL.icol);
      float* const a = base + sL.a;
      float* const Yh1 = base + sL.Yh1;
      float* const Yh2 = base + sL.Yh2;
      float* const zs = base + sL.zs + warp * 128;
      float* const bs = base + sL.bs;
      float* const W2s = base + sL.W2s;
      float* const q2 = base + sL.q2;
      const IdxT* const llist = reinterpret_cast<const IdxT*>(base + sL.llist);
      if (nlongF2 > 0) {
        for (int k = warp; k < nlongF2; k += nwarps)
          gather_long_row<IdxT, true>((int)llist[k], H4, lane, irp, icol, a, Yh1, HS, Yh2, HS, n, zs);
        __syncthreads();
      }
      for (int rb = warp * 32; rb < n1; rb += nwarps * 32) {
        const int re = min(rb + 32, n1);
        gather_short_rows<IdxT, true>(rb, re, H4, lane, irp, icol, a, Yh1, HS, Yh2, HS, n);
        __syncwarp();
        const int i = rb + lane;
        if (i < re) {
          float acc[HID];
#pragma unroll
          for (int c4 = 0; c4 < H4; ++c4) {
            const float4 b = ld4(bs + HID + 4 * c4);
            acc[4 * c4] = b.x; acc[4 * c4 + 1] = b.y; acc[4 * c4 + 2] = b.z; acc[4 * c4 + 3] = b.w;
          }
#pragma unroll
          for (int f4 = 0; f4 < H4; ++f4) {
            const float4 u = ld4(Yh2 + i * HS + 4 * f4);
            const float z[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float* wr = W2s + (4 * f4 + k) * HS;
#pragma unroll
              for (int c4 = 0; c4 < H4; ++c4) {
                const float4 w = ld4(wr + 4 * c4);
                acc[4 * c4] = fmaf(z[k], w.x, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(z[k], w.y, acc[4 * c4 + 1]);
                acc[4 * c4 + 2] = fmaf(z[k], w.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(z[k], w.w, acc[4 * c4 + 3]);
              }
            }
          }
          float ss = 0.f;
#pragma unroll
          for (int c = 0; c < HID; ++c) ss = fmaf(acc[c], acc[c], ss);
          const float q = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
          for (int c4 = 0; c4 < H4; ++c4)
            st4(Yh2 + i * HS + 4 * c4, make_float4(acc[4 * c4] / q, acc[4 * c4 + 1] / q, acc[4 * c4 + 2] / q, acc[4 * c4 + 3] / q));
          q2[i] = q;
        }
      }
      }
      __syncthreads();
      // ---- S: row r (= level-order id 0): layer 3, readout, softmax, -log p[gt], layer-3 backward
      if (warp == 0) {
        const IdxT* const irp = reinterpret_cast<const IdxT*>(base + sL.irp);
        const IdxT* const icol = reinterpret_cast<const IdxT*>(base + sL.icol);
        float* const a = base + sL.a;
        float* const Yh1 = base + sL.Yh1;
        float* const Yh2 = base + sL.Yh2;
        float* const zs = base + sL.zs + warp * 128;
        float* const bs = base + sL.bs;
        float* const W3s = base + sL.W3s;
        float* const logit = base + sL.logit;
        float* const dE = base + sL.dE;
        float* const dZ3 = base + sL.dZ3;
        const int r0 = irp[0], r1 = irp[1];
        float z = 0.f;
        if (lane < HID)
          for (int e = r0; e < r1; ++e) z = fmaf(a[e], fmaxf(Yh2[(int)icol[e] * HS + lane], 0.f), z);
        if (lane < HID) zs[lane] = z;
        __syncwarp();
        float y3 = lane < EMB ? bs[2 * HID + lane] : 0.f;
        if (lane < EMB)
          for (int f = 0; f < HID; ++f) y3 = fmaf(zs[f], W3s[f * EMB + lane], y3);
        const float ss = warp_sum(lane < EMB ? y3 * y3 : 0.f);
        const float q3 = fmaxf(sqrtf(ss), 1e-12f);
        const float yh3 = lane < EMB ? y3 / q3 : 0.f;
        const float e1v = lane < HID ? fmaxf(Yh1[lane], 0.f) : 0.f;  // row 0 of H1
        const float e2v = lane < HID ? fmaxf(Yh2[lane], 0.f) : 0.f;  // row 0 of H2
        // logits = pred_model(concat) (models.py:260,375), softmax over classes (explain.py:714)
        for (int c = 0; c < C; ++c) {
          const float* wp = m.Wp + c * PD;
          float t = 0.f;
          if (lane < HID) t = fmaf(e1v, __ldg(wp + lane), fmaf(e2v, __ldg(wp + HID + lane), t));
          if (lane < EMB) t = fmaf(yh3, __ldg(wp + 2 * HID + lane), t);
          t = warp_sum(t);
          if (lane == 0) logit[c] = t + __ldg(m.bp + c);
        }
        __syncwarp();
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, logit[c]);
        mx = warp_max(mx);
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += expf(logit[c] - mx);
        se = warp_sum(se);
        __syncwarp();
        for (int c = lane; c < C; c += 32)
          logit[c] = expf(logit[c] - mx) / se - (c == gt ? 1.f : 0.f);  // dL/dlogits = p - onehot(gt) (explain.py:750-753)
        __syncwarp();
        float d1 = 0.f, d2 = 0.f, d3 = 0.f;
        for (int c = 0; c < C; ++c) {
          const float gc = logit[c];
          const float* wp = m.Wp + c * PD;
          if (lane < HID) { d1 = fmaf(gc, __ldg(wp + lane), d1); d2 = fmaf(gc, __ldg(wp + HID + lane), d2); }
          if (lane < EMB) d3 = fmaf(gc, __ldg(wp + 2 * HID + lane), d3);
        }
        if (lane < HID) { dE[lane] = d1; dE[HS + lane] = d2; }
        // backward of y/max(|y|,eps): dY = (dYh - Yh <Yh,dYh>)/q ; dZ3 = dY3 W3^T
        const float s3 = warp_sum(yh3 * d3);
        const float dy3 = lane < EMB ? (d3 - yh3 * s3) / q3 : 0.f;
        __syncwarp();
        if (lane < EMB) zs[lane] = dy3;
        __syncwarp();
        if (lane < HID) dZ3[lane] = dot_v4(zs, W3s + lane * EMB, EMB / 4);
      }
      __syncthreads();
      // ---- B2: rows {r} U N(r) (one thread per row): dYh2 = dEmb2 (row r) + a[r,j] dZ3 (j in N(r)),
      //          relu', normalise', dZ2 = dY2 W2^T
      {
        const IdxT* const irp = reinterpret_cast<const IdxT*>(base + sL.irp);
        const IdxT* const icol = reinterpret_cast<const IdxT*>(base + sL.icol);
        float* const a = base + sL.a;
        float* const dE = base + sL.dE;
        float* const dZ3 = base + sL.dZ3;
        float* const Yh2 = base + sL.Yh2;
        float* const q2 = base + sL.q2;
        float* const W2s = base + sL.W2s;
        float* const dZ2 = base + sL.dZ2;
        const int r0 = irp[0];
        const int items = 1 + (int)irp[1] - r0;
        for (int item = tid; item < items; item += nthreads) {
          int j = 0;
          float coef = 1.f;
          const float* dsrc = dE + HS;
          if (item > 0) {
            const int e = r0 + item - 1;
            j = icol[e];
            coef = a[e];
            dsrc = dZ3;
          }
          const float* yr = Yh2 + j * HS;
          float dy[HID];
          float s = 0.f;
#pragma unroll
          for (int c4 = 0; c4 < H4; ++c4) {
            const float4 yh = ld4(yr + 4 * c4);
            const float4 g4 = ld4(dsrc + 4 * c4);
            dy[4 * c4] = yh.x > 0.f ? coef * g4.x : 0.f;      // relu backward: grad where input > 0
            dy[4 * c4 + 1] = yh.y > 0.f ? coef * g4.y : 0.f;
            dy[4 * c4 + 2] = yh.z > 0.f ? coef * g4.z : 0.f;
            dy[4 * c4 + 3] = yh.w > 0.f ? coef * g4.w : 0.f;
            s = fmaf(yh.x, dy[4 * c4], s); s = fmaf(yh.y, dy[4 * c4 + 1], s);
            s = fmaf(yh.z, dy[4 * c4 + 2], s); s = fmaf(yh.w, dy[4 * c4 + 3], s);
          }
          const float q = q2[j];
#pragma unroll
          for (int c4 = 0; c4 < H4; ++c4) {
            const float4 yh = ld4(yr + 4 * c4);
            dy[4 * c4] = (dy[4 * c4] - yh.x * s) / q; dy[4 * c4 + 1] = (dy[4 * c4 + 1] - yh.y * s) / q;
            dy[4 * c4 + 2] = (dy[4 * c4 + 2] - yh.z * s) / q; dy[4 * c4 + 3] = (dy[4 * c4 + 3] - yh.w * s) / q;
          }
#pragma unroll
          for (int f4 = 0; f4 < H4; ++f4) {
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float* wr = W2s + (4 * f4 + k) * HS;
              float t = 0.f;
#pragma unroll
              for (int c4 = 0; c4 < H4; ++c4) {
                const float4 w = ld4(wr + 4 * c4);
                t = fmaf(dy[4 * c4], w.x, t); t = fmaf(dy[4 * c4 + 1], w.y, t);
                t = fmaf(dy[4 * c4 + 2], w.z, t); t = fmaf(dy[4 * c4 + 3], w.w, t);
              }
              o[k] = t;
            }
            st4(dZ2 + j * HS + 4 * f4, make_float4(o[0], o[1], o[2], o[3]));
          }
        }
      }
      __syncthreads();
      // ---- B1: rows [0,n2): dH1 = A_m^T dZ2 (only columns < n1 carry gradient) lands in T[i]; then one
      //          thread per row: relu', normalise', dZ1 = dY1 W1^T, dL/dsF partial, T[i] = dZ1 (.) sF
      {
      const IdxT* const irp = reinterpret_cast<const IdxT*>(base + sL.irp);
      const IdxT* const icol = reinterpret_cast<const IdxT*>(base + sL.icol);
      float* const a = base + sL.a;
      float* const dZ2 = base + sL.dZ2;
      float* const Tb = base + sL.T;
      float* const Yh1 = base + sL.Yh1;
      float* const dE = base + sL.dE;
      float* const q1 = base + sL.q1;
      float* const W1s = base + sL.W1s;
      float* const U = base + sL.U;
      float* const sF = base + sL.sF;
      float* const gFp = base + sL.gFp;
      for (int rb = warp * 32; rb < n2; rb += nwarps * 32) {
        const int re = min(rb + 32, n2);
        {  // all rows by lane groups: at most n1 columns per row carry gradient
          const int grp = lane / H4, q = lane - grp * H4;
          const int epi = 32 / H4;
          if (grp < epi) {
            for (int i = rb + grp; i < re; i += epi) {
              const int r0 = irp[i], r1 = irp[i + 1];
              float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
              for (int e = r0; e < r1; ++e) {
                const int c = icol[e];
                if (c >= n1) break;
                fma4(acc, a[e], ld4(dZ2 + c * HS + 4 * q));
              }
              st4(Tb + i * TS + 4 * q, acc);
            }
          }
        }
        __syncwarp();
        const int i = rb + lane;
        const bool valid = i < re;
        float dy[HID];
        {
          const float* yr = Yh1 + (valid ? i : 0) * HS;
          const float* tr = Tb + (valid ? i : 0) * TS;
          float s = 0.f;
#pragma unroll
          for (int c4 = 0; c4 < H4; ++c4) {
            const float4 yh = ld4(yr + 4 * c4);
            float4 g4 = ld4(tr + 4 * c4);
            if (i == 0) { const float4 e4 = ld4(dE + 4 * c4); g4.x += e4.x; g4.y += e4.y; g4.z += e4.z; g4.w += e4.w; }
            dy[4 * c4] = yh.x > 0.f ? g4.x : 0.f; dy[4 * c4 + 1] = yh.y > 0.f ? g4.y : 0.f;
            dy[4 * c4 + 2] = yh.z > 0.f ? g4.z : 0.f; dy[4 * c4 + 3] = yh.w > 0.f ? g4.w : 0.f;
            s = fmaf(yh.x, dy[4 * c4], s); s = fmaf(yh.y, dy[4 * c4 + 1], s);
            s = fmaf(yh.z, dy[4 * c4 + 2], s); s = fmaf(yh.w, dy[4 * c4 + 3], s);
          }
          const float q = valid ? q1[i] : 1.f;
#pragma unroll
          for (int c4 = 0; c4 < H4; ++c4) {
            const float4 yh = ld4(yr + 4 * c4);
            dy[4 * c4] = (dy[4 * c4] - yh.x * s) / q; dy[4 * c4 + 1] = (dy[4 * c4 + 1] - yh.y * s) / q;
            dy[4 * c4 + 2] = (dy[4 * c4 + 2] - yh.z * s) / q; dy[4 * c4 + 3] = (dy[4 * c4 + 3] - yh.w * s) / q;
          }
        }
        for (int f4 = 0; f4 < D4; ++f4) {
          float o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float* wr = W1s + (4 * f4 + k) * HS;
            float t = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < H4; ++c4) {
              const float4 w = ld4(wr + 4 * c4);
              t = fmaf(dy[4 * c4], w.x, t); t = fmaf(dy[4 * c4 + 1], w.y, t);
              t = fmaf(dy[4 * c4 + 2], w.z, t); t = fmaf(dy[4 * c4 + 3], w.w, t);
            }
            o[k] = t;
          }
          float4 pu = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid) {
            const float4 u = ld4(U + i * dp + 4 * f4);
            const float4 s4 = ld4(sF + 4 * f4);
            pu = make_float4(o[0] * u.x, o[1] * u.y, o[2] * u.z, o[3] * u.w);
            st4(Tb + i * TS + 4 * f4, make_float4(o[0] * s4.x, o[1] * s4.y, o[2] * s4.z, o[3] * s4.w));
          }
          pu.x = warp_sum(pu.x); pu.y = warp_sum(pu.y); pu.z = warp_sum(pu.z); pu.w = warp_sum(pu.w);
          if (lane == 0) {
            float4 g4 = ld4(gFp + warp * dp + 4 * f4);
            g4.x += pu.x; g4.y += pu.y; g4.z += pu.z; g4.w += pu.w;
            st4(gFp + warp * dp + 4 * f4, g4);
          }
        }
      }
      }
      __syncthreads();
      // ---- P: per undirected edge: dA_ij, dA_ji, symmetrise, regularisers, Adam, next mask value
      {
        float* const gFp = base + sL.gFp;
        float* const sF = base + sL.sF;
        float* const Fm = base + sL.F; float* const mF = base + sL.mF; float* const vF = base + sL.vF;
        const IdxT* const pi = reinterpret_cast<const IdxT*>(base + sL.pi); const IdxT* const pj = reinterpret_cast<const IdxT*>(base + sL.pj);
        const IdxT* const ppij = reinterpret_cast<const IdxT*>(base + sL.ppij); const IdxT* const ppji = reinterpret_cast<const IdxT*>(base + sL.ppji);
        float* const lap2 = base + sL.lap2;
        float* const Tb = base + sL.T;
        float* const X = base + sL.X;
        float* const dZ2 = base + sL.dZ2;
        float* const Yh1 = base + sL.Yh1;
        float* const dZ3 = base + sL.dZ3;
        float* const Yh2 = base + sL.Yh2;
        float2* const MM = reinterpret_cast<float2*>(base + sL.M); float2* const mm = MM + np; float2* const vv = mm + np; float2* const SS = vv + np;
        float* const a = base + sL.a;
        const float2 tab = __ldg(hp.adam_tab + (it - 1));
        const float step = tab.x, bc2s = tab.y;
        const bool last = (it == hp.iters);
        // feature mask: dL/dF = sF(1-sF) (sum_i dZ1[i] U[i] + feat_size/d) ; Adam (explain.py:766, train_utils.py:10)
        for (int f = tid; f < d; f += nthreads) {
          float gsum = 0.f;
          for (int w = 0; w < nwarps; ++w) { gsum += gFp[w * dp + f]; gFp[w * dp + f] = 0.f; }
          const float s = sF[f];
          const float g = s * (1.f - s) * (gsum + hp.c_feat_size / (float)d);
          float mf = mF[f], vf = vF[f], Fv = Fm[f];
          mf = mf + (g - mf) * hp.one_minus_b1;
          vf = vf * hp.b2 + hp.one_minus_b2 * g * g;
          Fv = Fv - step * (mf / (sqrtf(vf) / bc2s + hp.eps));
          mF[f] = mf; vF[f] = vf; Fm[f] = Fv;
          sF[f] = sigmoid_f(Fv);
        }
        for (int p = tid; p < np; p += nthreads) {
          const int i = pi[p], j = pj[p];
          float G = lap2[p];
          if (i < n2) G += dot_v4(Tb + i * TS, X + j * dp, D4);
          if (j < n2) G += dot_v4(Tb + j * TS, X + i * dp, D4);
          if (i < n1) G += dot_relu_v4(dZ2 + i * HS, Yh1 + j * HS, H4);
          if (j < n1) G += dot_relu_v4(dZ2 + j * HS, Yh1 + i * HS, H4);
          if (i == 0) G += dot_relu_v4(dZ3, Yh2 + j * HS, H4);
          G *= 0.5f;  // sym_mask = (S + S^T)/2 (explain.py:671)
          float2 Mv = MM[p];
          const float2 Sv = SS[p];
          // size: coeff*sum(S) ; entropy: mean over n^2 of H(S), dH/dM = -M S(1-S) (explain.py:755-770)
          const float gi = Sv.x * (1.f - Sv.x) * (G + hp.c_size - ent_over_nn * Mv.x);
          const float gj = Sv.y * (1.f - Sv.y) * (G + hp.c_size - ent_over_nn * Mv.y);
          float2 m2 = mm[p], v2 = vv[p];
          m2.x = m2.x + (gi - m2.x) * hp.one_minus_b1;
          m2.y = m2.y + (gj - m2.y) * hp.one_minus_b1;
          v2.x = v2.x * hp.b2 + hp.one_minus_b2 * gi * gi;
          v2.y = v2.y * hp.b2 + hp.one_minus_b2 * gj * gj;
          Mv.x = Mv.x - step * (m2.x / (sqrtf(v2.x) / bc2s + hp.eps));
          Mv.y = Mv.y - step * (m2.y / (sqrtf(v2.y) / bc2s + hp.eps));
          const float2 Sn = make_float2(sigmoid_f(Mv.x), sigmoid_f(Mv.y));
          MM[p] = Mv; mm[p] = m2; vv[p] = v2; SS[p] = Sn;
          const float an = 0.5f * (Sn.x + Sn.y);
          const IdxT pa = ppij[p], pb = ppji[p];
          if (pa != kNone) a[pa] = an;
          if (pb != kNone) a[pb] = an;
          if (last) {
            A.out_mask[edge_off + A.plan.pair_oij[pair_off + p]] = an;
            A.out_mask[edge_off + A.plan.pair_oji[pair_off + p]] = an;
          }
        }
      }
      __syncthreads();
    }
    if (A.out_feat != nullptr) {
      const float* const sF = base + sL.sF;
      for (int f = tid; f < d; f += nthreads) A.out_feat[(int64_t)task_id * d + f] = sF[f];
    }
    __syncthreads();
  }
}

template <bool kShared, typename IdxT, int HID, int EMB, int NT>
cudaError_t launch_one(const GxExplainLaunch& cfg, const ExplainArgs& args, cudaStream_t s) {
  auto kern = explain_node_kernel<kShared, IdxT, HID, EMB, NT>;
  if (kShared) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg.smem_bytes);
    if (e != cudaSuccess) return e;
  }
  kern<<<cfg.grid, cfg.threads, kShared ? cfg.smem_bytes : 0, s>>>(args);
  return cudaGetLastError();
}

template <int HID, int EMB>
cudaError_t launch_dims(const GxExplainLaunch& cfg, const ExplainArgs& args, cudaStream_t s) {
  if (cfg.smem_bytes > 0) {
    if (cfg.threads <= 256) return launch_one<true, uint16_t, HID, EMB, 256>(cfg, args, s);
    return launch_one<true, uint16_t, HID, EMB, 512>(cfg, args, s);
  }
  return launch_one<false, int32_t, HID, EMB, 512>(cfg, args, s);
}

}  // namespace

int gx_explain_max_smem() { return 227 * 1024; }

cudaError_t gx_launch_explain(const GxExplainLaunch& cfg, const GxGraphDev& g, const GxModelDev& m,
                              const GxHparamsDev& hp, const GxPlanArrays& plan, const float* m0,
                              float* out_mask, float* out_feat, cudaStream_t s) {
  ExplainArgs args;
  args.order = cfg.order; args.ntasks = cfg.ntasks; args.counter = cfg.counter;
  args.gws = cfg.gws; args.gws_stride_words = cfg.gws_stride_words;
  args.g = g; args.m = m; args.hp = hp; args.plan = plan;
  args.m0 = m0; args.out_mask = out_mask; args.out_feat = out_feat;
  if (m.hid == 20 && m.emb == 20) return launch_dims<20, 20>(cfg, args, s);
  return cudaErrorInvalidValue;  // gx_set_model rejects other widths before this point
}
