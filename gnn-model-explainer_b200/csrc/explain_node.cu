// explain_node.cu -- K2: the persistent per-node mask-optimisation kernel (node mode).
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
//     of the explained node; with nodes relabelled in (distance, id) order every layer's row
//     set is a prefix [0,n_{L-l}) and the backward touches the same prefixes.
//   * dL/dF needs sum_i dZ1[i] (.) U[i] with U = A_m X kept from the forward, so the layer-1
//     transpose aggregation disappears.
//   * The returned mask is the one built in the forward of the LAST epoch (explain.py:694,209),
//     i.e. after num_epochs-1 updates; the last backward/Adam step is unobservable and skipped.
//
// Phases per epoch (one __syncthreads each): F1 | F2 | S (row r: layer 3 + readout + softmax +
// layer-3 backward, one warp) | B2 | B1 | P (per undirected edge: SDDMM dots, symmetrise,
// regularisers, Adam on both directions, next epoch's mask value).
#include "gnnx_internal.cuh"

namespace {

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

// dot of two length-(4*n4) shared vectors
__device__ __forceinline__ float dot_v4(const float* __restrict__ a, const float* __restrict__ b, int n4) {
  float s = 0.f;
  for (int k = 0; k < n4; ++k) {
    const float4 x = reinterpret_cast<const float4*>(a)[k];
    const float4 y = reinterpret_cast<const float4*>(b)[k];
    s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
  }
  return s;
}
// dot(a, relu(b))
__device__ __forceinline__ float dot_relu_v4(const float* __restrict__ a, const float* __restrict__ b, int n4) {
  float s = 0.f;
  for (int k = 0; k < n4; ++k) {
    const float4 x = reinterpret_cast<const float4*>(a)[k];
    const float4 y = reinterpret_cast<const float4*>(b)[k];
    s = fmaf(x.x, fmaxf(y.x, 0.f), s); s = fmaf(x.y, fmaxf(y.y, 0.f), s);
    s = fmaf(x.z, fmaxf(y.z, 0.f), s); s = fmaf(x.w, fmaxf(y.w, 0.f), s);
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

// y[lane] = sum_f zs[f] * w[f]  with w in registers (one column / row of a small dense matrix)
template <int K>
__device__ __forceinline__ float dense_reg(const float* __restrict__ zs, const float (&w)[K], float acc) {
#pragma unroll
  for (int f4 = 0; f4 < K / 4; ++f4) {
    const float4 z = reinterpret_cast<const float4*>(zs)[f4];
    acc = fmaf(z.x, w[4 * f4 + 0], acc); acc = fmaf(z.y, w[4 * f4 + 1], acc);
    acc = fmaf(z.z, w[4 * f4 + 2], acc); acc = fmaf(z.w, w[4 * f4 + 3], acc);
  }
#pragma unroll
  for (int f = K / 4 * 4; f < K; ++f) acc = fmaf(zs[f], w[f], acc);
  return acc;
}

template <bool kShared, typename IdxT, int HID, int EMB, int DCH, int NT>
__global__ void __launch_bounds__(NT) explain_node_kernel(const ExplainArgs A) {
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ int s_task;
  constexpr IdxT kNone = IdxTraits<IdxT>::kNone;
  constexpr int HS = (HID + 3) / 4 * 4;
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
    const GxLayout L = gx_make_layout(n, n1, n2, e1, np, d, HID, EMB, C, nwarps, (int)sizeof(IdxT));
    const int dp = L.dp;
    float* const X = base + L.X;
    float* const U = base + L.U;
    float* const Yh1 = base + L.Yh1;
    float* const q1 = base + L.q1;
    float* const Yh2 = base + L.Yh2;
    float* const q2 = base + L.q2;
    float* const dZ2 = base + L.dZ2;
    float* const dZ1s = base + L.dZ1s;
    float* const a = base + L.a;
    float* const Mij = base + L.M;
    float* const Mji = Mij + np;
    float* const mij = Mji + np;
    float* const mji = mij + np;
    float* const vij = mji + np;
    float* const vji = vij + np;
    float* const lap2 = base + L.lap2;
    float* const W1s = base + L.W1s;
    float* const sF = base + L.sF;
    float* const Fm = base + L.F;
    float* const mF = base + L.mF;
    float* const vF = base + L.vF;
    float* const gFp = base + L.gFp;
    int zw = dp > HS ? dp : HS;
    zw = zw > ((EMB + 3) / 4 * 4) ? zw : ((EMB + 3) / 4 * 4);
    float* const zs = base + L.zs + warp * zw;
    float* const dE = base + L.dE;
    float* const dZ3 = base + L.dZ3;
    float* const logit = base + L.logit;
    IdxT* const icol = reinterpret_cast<IdxT*>(base + L.icol);
    IdxT* const irp = reinterpret_cast<IdxT*>(base + L.irp);
    IdxT* const pi = reinterpret_cast<IdxT*>(base + L.pi);
    IdxT* const pj = reinterpret_cast<IdxT*>(base + L.pj);
    IdxT* const ppij = reinterpret_cast<IdxT*>(base + L.ppij);
    IdxT* const ppji = reinterpret_cast<IdxT*>(base + L.ppji);
    const int32_t* __restrict__ lo2gid = A.plan.lo2gid + node_off;
    const float nn = (float)n * (float)n;
    const float ent_over_nn = hp.c_ent / nn;
    const float lap_over_nn = hp.c_lap / nn;

    // ------------------------------------------------------------------ load
    for (int idx = tid; idx < n * dp; idx += nthreads) {
      const int i = idx / dp, f = idx - i * dp;
      X[idx] = f < d ? __ldg(A.g.feat + (int64_t)lo2gid[i] * d + f) : 0.f;
    }
    for (int idx = tid; idx < d * HS; idx += nthreads) {
      const int f = idx / HS, c = idx - f * HS;
      W1s[idx] = c < HID ? __ldg(m.W[0] + f * HID + c) : 0.f;
    }
    for (int e = tid; e < e1; e += nthreads) icol[e] = (IdxT)A.plan.icol[edge_off + e];
    for (int i = tid; i <= n2; i += nthreads) irp[i] = (IdxT)A.plan.irowptr[rp_off + i];
    for (int f = tid; f < dp; f += nthreads) {
      sF[f] = 0.5f;  // sigmoid(0): feat_mask is initialised to 0 (explain.py:633-643)
      Fm[f] = 0.f; mF[f] = 0.f; vF[f] = 0.f;
    }
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
      Mij[p] = Mi; Mji[p] = Mj;
      mij[p] = 0.f; mji[p] = 0.f; vij[p] = 0.f; vji[p] = 0.f;
      const float yi = (float)__ldg(A.g.pred_label + lo2gid[i]);
      const float yj = (float)__ldg(A.g.pred_label + lo2gid[j]);
      lap2[p] = lap_over_nn * (yi - yj) * (yi - yj);  // d/dA_ij + d/dA_ji of y^T (D - A) y / n^2
      const float a0 = 0.5f * (sigmoid_f(Mi) + sigmoid_f(Mj));  // explain.py:665-678
      if (i < n2) a[pij] = a0;
      if (j < n2) a[pji] = a0;
      if (hp.iters == 0) {
        A.out_mask[edge_off + oij] = a0;
        A.out_mask[edge_off + oji] = a0;
      }
    }
    __syncthreads();

    // ------------------------------------------------------------------ epochs
    for (int it = 1; it <= hp.iters; ++it) {
      // ---- F1: rows [0,n2): U = A_m X ; Y1 = (U . sF) W1 + b1 ; normalise           (models.py:70-78)
      {
        float sFl[DCH];
#pragma unroll
        for (int ch = 0; ch < DCH; ++ch) sFl[ch] = (ch * 32 + lane < dp) ? sF[ch * 32 + lane] : 0.f;
        const float b1 = lane < HID ? __ldg(m.b[0] + lane) : 0.f;
        for (int i = warp; i < n2; i += nwarps) {
          const int r0 = irp[i], r1 = irp[i + 1];
          float u[DCH];
#pragma unroll
          for (int ch = 0; ch < DCH; ++ch) u[ch] = 0.f;
#pragma unroll 4
          for (int e = r0; e < r1; ++e) {
            const int c = icol[e];
            const float av = a[e];
            const float* xr = X + c * dp;
#pragma unroll
            for (int ch = 0; ch < DCH; ++ch)
              if (ch * 32 + lane < dp) u[ch] = fmaf(av, xr[ch * 32 + lane], u[ch]);
          }
#pragma unroll
          for (int ch = 0; ch < DCH; ++ch)
            if (ch * 32 + lane < dp) {
              U[i * dp + ch * 32 + lane] = u[ch];
              zs[ch * 32 + lane] = u[ch] * sFl[ch];  // x * sigmoid(feat_mask) (explain.py:707), linear in x
            }
          __syncwarp();
          float y = b1;
          if (lane < HS) {
            for (int f = 0; f < d; ++f) y = fmaf(zs[f], W1s[f * HS + lane], y);
          }
          const float ss = warp_sum(lane < HID ? y * y : 0.f);
          const float q = fmaxf(sqrtf(ss), 1e-12f);  // F.normalize(p=2, dim=2), eps 1e-12
          if (lane < HS) Yh1[i * HS + lane] = lane < HID ? y / q : 0.f;
          if (lane == 0) q1[i] = q;
          __syncwarp();
        }
      }
      __syncthreads();
      // ---- F2: rows [0,n1): Y2 = (A_m relu(Yh1)) W2 + b2 ; normalise
      {
        float wcol[HID];
#pragma unroll
        for (int f = 0; f < HID; ++f) wcol[f] = lane < HID ? __ldg(m.W[1] + f * HID + lane) : 0.f;
        const float b2 = lane < HID ? __ldg(m.b[1] + lane) : 0.f;
        for (int i = warp; i < n1; i += nwarps) {
          const int r0 = irp[i], r1 = irp[i + 1];
          float z = 0.f;
#pragma unroll 4
          for (int e = r0; e < r1; ++e) {
            const int c = icol[e];
            const float av = a[e];
            if (lane < HS) z = fmaf(av, fmaxf(Yh1[c * HS + lane], 0.f), z);
          }
          if (lane < HS) zs[lane] = z;
          __syncwarp();
          const float y = dense_reg<HID>(zs, wcol, b2);
          const float ss = warp_sum(lane < HID ? y * y : 0.f);
          const float q = fmaxf(sqrtf(ss), 1e-12f);
          if (lane < HS) Yh2[i * HS + lane] = lane < HID ? y / q : 0.f;
          if (lane == 0) q2[i] = q;
          __syncwarp();
        }
      }
      __syncthreads();
      // ---- S: row r (= level-order id 0): layer 3, readout, softmax, -log p[gt], layer-3 backward
      if (warp == 0) {
        const int r0 = irp[0], r1 = irp[1];
        float z = 0.f;
        for (int e = r0; e < r1; ++e) {
          const int c = icol[e];
          if (lane < HS) z = fmaf(a[e], fmaxf(Yh2[c * HS + lane], 0.f), z);
        }
        if (lane < HS) zs[lane] = z;
        __syncwarp();
        float y3 = lane < EMB ? __ldg(m.b[2] + lane) : 0.f;
        if (lane < EMB)
          for (int f = 0; f < HID; ++f) y3 = fmaf(zs[f], __ldg(m.W[2] + f * EMB + lane), y3);
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
        if (lane < HS) { dE[lane] = lane < HID ? d1 : 0.f; dE[HS + lane] = lane < HID ? d2 : 0.f; }
        // backward of y/max(|y|,eps): dY = (dYh - Yh <Yh,dYh>)/q ; dZ3 = dY3 W3^T
        const float s3 = warp_sum(yh3 * d3);
        const float dy3 = lane < EMB ? (d3 - yh3 * s3) / q3 : 0.f;
        __syncwarp();
        if (lane < ((EMB + 3) / 4 * 4)) zs[lane] = dy3;
        __syncwarp();
        float dz = 0.f;
        if (lane < HID)
          for (int c = 0; c < EMB; ++c) dz = fmaf(zs[c], __ldg(m.Wt[2] + c * HID + lane), dz);
        if (lane < HS) dZ3[lane] = lane < HID ? dz : 0.f;
      }
      __syncthreads();
      // ---- B2: rows {r} U N(r): dYh2 = dEmb2 (row r) + a[r,j] dZ3 (j in N(r)), relu', normalise', W2^T
      {
        float wrow[HID];
#pragma unroll
        for (int c = 0; c < HID; ++c) wrow[c] = lane < HID ? __ldg(m.Wt[1] + c * HID + lane) : 0.f;
        const int r0 = irp[0];
        const int items = 1 + (int)irp[1] - r0;
        const float dz3 = lane < HS ? dZ3[lane] : 0.f;
        const float de2 = lane < HS ? dE[HS + lane] : 0.f;
        for (int item = warp; item < items; item += nwarps) {
          int j = 0;
          float dyh = de2;
          if (item > 0) {
            const int e = r0 + item - 1;
            j = icol[e];
            dyh = a[e] * dz3;
          }
          const float yh = lane < HS ? Yh2[j * HS + lane] : 0.f;
          if (!(yh > 0.f)) dyh = 0.f;  // relu backward: grad where input > 0
          const float s = warp_sum(yh * dyh);
          const float dy = (dyh - yh * s) / q2[j];
          if (lane < HS) zs[lane] = dy;
          __syncwarp();
          const float dz = dense_reg<HID>(zs, wrow, 0.f);
          if (lane < HS) dZ2[j * HS + lane] = lane < HID ? dz : 0.f;
          __syncwarp();
        }
      }
      __syncthreads();
      // ---- B1: rows [0,n2): dH1 = A_m^T dZ2 (only columns < n1 carry gradient), relu', normalise', W1^T,
      //          dL/dsF accumulation, dZ1 (.) sF kept for the edge dots
      {
        float gacc[DCH];
        float sFl[DCH];
#pragma unroll
        for (int ch = 0; ch < DCH; ++ch) {
          gacc[ch] = 0.f;
          sFl[ch] = (ch * 32 + lane < dp) ? sF[ch * 32 + lane] : 0.f;
        }
        const float de1 = lane < HS ? dE[lane] : 0.f;
        for (int i = warp; i < n2; i += nwarps) {
          const int r0 = irp[i], r1 = irp[i + 1];
          float dh = (i == 0) ? de1 : 0.f;
          for (int e = r0; e < r1; ++e) {
            const int c = icol[e];
            if (c >= n1) break;  // columns ascend in level order: the rest has no dZ2
            if (lane < HS) dh = fmaf(a[e], dZ2[c * HS + lane], dh);
          }
          const float yh = lane < HS ? Yh1[i * HS + lane] : 0.f;
          const float dyh = (yh > 0.f) ? dh : 0.f;
          const float s = warp_sum(yh * dyh);
          const float dy = (dyh - yh * s) / q1[i];
          if (lane < HS) zs[lane] = dy;
          __syncwarp();
#pragma unroll
          for (int ch = 0; ch < DCH; ++ch) {
            const int f = ch * 32 + lane;
            if (f < d) {
              const float dz = dot_v4(zs, W1s + f * HS, HS / 4);
              gacc[ch] = fmaf(dz, U[i * dp + f], gacc[ch]);
              dZ1s[i * dp + f] = dz * sFl[ch];
            } else if (f < dp) {
              dZ1s[i * dp + f] = 0.f;
            }
          }
          __syncwarp();
        }
#pragma unroll
        for (int ch = 0; ch < DCH; ++ch)
          if (ch * 32 + lane < dp) gFp[warp * dp + ch * 32 + lane] = gacc[ch];
      }
      __syncthreads();
      // ---- P: per undirected edge: dA_ij, dA_ji, symmetrise, regularisers, Adam, next mask value
      {
        const float2 tab = __ldg(hp.adam_tab + (it - 1));
        const float step = tab.x, bc2s = tab.y;
        const bool last = (it == hp.iters);
        // feature mask: dL/dF = sF(1-sF) (sum_i dZ1[i] U[i] + feat_size/d) ; Adam (explain.py:766, train_utils.py:10)
        for (int f = tid; f < d; f += nthreads) {
          float gsum = 0.f;
          for (int w = 0; w < nwarps; ++w) gsum += gFp[w * dp + f];
          const float s = sF[f];
          const float g = s * (1.f - s) * (gsum + hp.c_feat_size / (float)d);
          float mm = mF[f], vv = vF[f], Fv = Fm[f];
          mm = mm + (g - mm) * hp.one_minus_b1;
          vv = vv * hp.b2 + hp.one_minus_b2 * g * g;
          Fv = Fv - step * (mm / (sqrtf(vv) / bc2s + hp.eps));
          mF[f] = mm; vF[f] = vv; Fm[f] = Fv;
          sF[f] = sigmoid_f(Fv);
        }
        for (int p = tid; p < np; p += nthreads) {
          const int i = pi[p], j = pj[p];
          float G = lap2[p];
          if (i < n2) G += dot_v4(dZ1s + i * dp, X + j * dp, dp / 4);
          if (j < n2) G += dot_v4(dZ1s + j * dp, X + i * dp, dp / 4);
          if (i < n1) G += dot_relu_v4(dZ2 + i * HS, Yh1 + j * HS, HS / 4);
          if (j < n1) G += dot_relu_v4(dZ2 + j * HS, Yh1 + i * HS, HS / 4);
          if (i == 0) G += dot_relu_v4(dZ3, Yh2 + j * HS, HS / 4);
          G *= 0.5f;  // sym_mask = (S + S^T)/2 (explain.py:671)
          float Mi = Mij[p], Mj = Mji[p];
          const float Si = sigmoid_f(Mi), Sj = sigmoid_f(Mj);
          // size: coeff*sum(S) ; entropy: mean over n^2 of H(S), dH/dM = -M S(1-S) (explain.py:755-770)
          const float gi = Si * (1.f - Si) * (G + hp.c_size - ent_over_nn * Mi);
          const float gj = Sj * (1.f - Sj) * (G + hp.c_size - ent_over_nn * Mj);
          float mi_ = mij[p], mj_ = mji[p], vi_ = vij[p], vj_ = vji[p];
          mi_ = mi_ + (gi - mi_) * hp.one_minus_b1;
          mj_ = mj_ + (gj - mj_) * hp.one_minus_b1;
          vi_ = vi_ * hp.b2 + hp.one_minus_b2 * gi * gi;
          vj_ = vj_ * hp.b2 + hp.one_minus_b2 * gj * gj;
          Mi = Mi - step * (mi_ / (sqrtf(vi_) / bc2s + hp.eps));
          Mj = Mj - step * (mj_ / (sqrtf(vj_) / bc2s + hp.eps));
          Mij[p] = Mi; Mji[p] = Mj; mij[p] = mi_; mji[p] = mj_; vij[p] = vi_; vji[p] = vj_;
          const float an = 0.5f * (sigmoid_f(Mi) + sigmoid_f(Mj));
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
    if (A.out_feat != nullptr)
      for (int f = tid; f < d; f += nthreads) A.out_feat[(int64_t)task_id * d + f] = sF[f];
    __syncthreads();
  }
}

template <bool kShared, typename IdxT, int HID, int EMB, int DCH, int NT>
cudaError_t launch_one(const GxExplainLaunch& cfg, const ExplainArgs& args, cudaStream_t s) {
  auto kern = explain_node_kernel<kShared, IdxT, HID, EMB, DCH, NT>;
  if (kShared) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg.smem_bytes);
    if (e != cudaSuccess) return e;
  }
  kern<<<cfg.grid, cfg.threads, kShared ? cfg.smem_bytes : 0, s>>>(args);
  return cudaGetLastError();
}

template <int HID, int EMB, int DCH>
cudaError_t launch_dims(const GxExplainLaunch& cfg, const ExplainArgs& args, cudaStream_t s) {
  if (cfg.smem_bytes > 0) {
    if (cfg.idx16) {
      if (cfg.threads <= 256) return launch_one<true, uint16_t, HID, EMB, DCH, 256>(cfg, args, s);
      return launch_one<true, uint16_t, HID, EMB, DCH, 512>(cfg, args, s);
    }
    return launch_one<true, int32_t, HID, EMB, DCH, 512>(cfg, args, s);
  }
  return launch_one<false, int32_t, HID, EMB, DCH, 512>(cfg, args, s);
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
  const bool d1 = m.d <= 32;
  if (m.hid == 20 && m.emb == 20)
    return d1 ? launch_dims<20, 20, 1>(cfg, args, s) : launch_dims<20, 20, 4>(cfg, args, s);
  return cudaErrorInvalidValue;  // host code pads other widths before calling (see api.cu)
}
