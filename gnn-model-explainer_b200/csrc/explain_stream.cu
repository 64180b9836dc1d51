// explain_stream.cu -- K2s: the mask-optimisation kernel for explained nodes whose k-hop state does not fit the
// 227 KB of shared memory (wide features and/or 10^4..10^5-node neighbourhoods: BASELINE config 5).
//
// Same arithmetic contract as explain_node.cu (explainer/explain.py:137-146,665-715,740-808 + autograd + Adam,
// models.py:58-80,230-267,363-376), different data placement and a different contraction order:
//   * one persistent CTA (1024 threads) per task, model weights and per-warp scratch in shared memory, every
//     per-node / per-edge array in a per-CTA global slab (L2 / HBM), the CSR and pair index arrays read in
//     place from the plan (no per-task copy);
//   * layer 1 is evaluated as A_m (X' W1) instead of (A_m X') W1: the d-wide feature row of a node is read
//     once per pass (F0: P = (X . sF) W1 for all nodes; B0: dL/dsF = sum_j X_j . (dP_j W1^T)) and every
//     per-edge operation is hid-wide (gathers of 80-byte rows, 20-float dots) -- for d = 128 that is 6.4x fewer
//     bytes per edge than the U = A_m X order the shared-memory kernel uses for d <= hid;
//   * dP = A_m^T dY1 needs, for every node j (also the outermost ones), its neighbours inside the layer-1 row
//     set: the plan's level-partitioned rows give that as a prefix of row j (cnt2).
// Phases per epoch (one __syncthreads each): F0 | F1 | F2 | S | B2 | B1 | B0 | P.
// Pairs between two outermost nodes are regulariser-only scalar recurrences (outer_pairs_kernel).
#include "explain_common.cuh"

namespace {

struct StreamSmem {
  int W1s, W1t, W2s, W2t, W3s, bs, sF, F, mF, vF, zs, dE, dZ3, logit, Wp, total;
};
__host__ __device__ inline StreamSmem stream_smem(int dp, int hid, int emb, int C, int nwarps) {
  StreamSmem S;
  int o = 0;
  auto take = [&](int words) { int r = o; o += gx_round_up(words, 4); return r; };
  S.W1s = take(dp * hid); S.W1t = take(hid * dp); S.W2s = take(hid * hid); S.W2t = take(hid * hid);
  S.W3s = take(hid * emb); S.bs = take(2 * hid + emb);
  S.sF = take(dp); S.F = take(dp); S.mF = take(dp); S.vF = take(dp);
  S.zs = take(nwarps * 128);
  S.dE = take(2 * hid); S.dZ3 = take(hid); S.logit = take(C < 32 ? 32 : C);
  S.Wp = take(C * (2 * hid + emb + 1) <= GX_WP_SMEM_MAX ? C * (2 * hid + emb + 1) : 0);
  S.total = o;
  return S;
}

// this lane's float4 (features 4q..4q+3) of a feature row of the full graph
__device__ __forceinline__ float4 load_x4(const float* __restrict__ row, int q, int d, bool vec) {
  if (vec) return __ldg(reinterpret_cast<const float4*>(row) + q);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  const int f = 4 * q;
  if (f < d) v.x = __ldg(row + f);
  if (f + 1 < d) v.y = __ldg(row + f + 1);
  if (f + 2 < d) v.z = __ldg(row + f + 2);
  if (f + 3 < d) v.w = __ldg(row + f + 3);
  return v;
}

// first slot in [r0,r1) whose column is >= bound (columns are partitioned by level, so the predicate is monotone)
__device__ __forceinline__ int prefix_below(const int32_t* __restrict__ icol, int r0, int r1, int bound) {
  int lo = r0, hi = r1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(icol + mid) < bound) lo = mid + 1; else hi = mid;
  }
  return lo - r0;
}

// warp w builds the ascending list of rows i < R with len(i) > kLongRow; returns the count (and, in *below,
// how many of them are < split)
template <typename LenF>
__device__ __forceinline__ int build_long_list(int R, int split, int32_t* list, int lane, LenF len, int* below) {
  int cnt = 0, cb = 0;
  for (int b0 = 0; b0 < R; b0 += 32) {
    const int i = b0 + lane;
    const bool lg = i < R && len(i) > kLongRow;
    const uint32_t bal = __ballot_sync(0xffffffffu, lg);
    if (lg) list[cnt + __popc(bal & ((1u << lane) - 1u))] = i;
    cnt += __popc(bal);
    cb += __popc(__ballot_sync(0xffffffffu, lg && i < split));
  }
  *below = cb;
  return cnt;
}

template <int HID, int EMB, int NT>
__global__ void __launch_bounds__(NT, 1) explain_stream_kernel(const ExplainArgs A) {
  extern __shared__ __align__(16) float sm[];
  __shared__ int s_task;
  __shared__ int s_long[4];  // long rows: among [0,n2), among [0,n1), rows < n2 with a long < n1 prefix, rows < n with a long < n2 prefix
  static_assert(HID % 4 == 0 && EMB % 4 == 0 && HID <= 32 && EMB <= 32, "hidden widths: multiples of 4, <= 32");
  constexpr int HS = HID, H4 = HID / 4, PD = 2 * HID + EMB;
  constexpr int nwarps = NT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const GxModelDev& m = A.m;
  const GxHparamsDev& hp = A.hp;
  const int d = m.d, C = m.C;
  const int dp = gx_round_up(d, 4), D4 = dp / 4;
  const bool xvec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(A.g.feat) & 15) == 0);
  const StreamSmem S = stream_smem(dp, HID, EMB, C, nwarps);
  float* const W1s = sm + S.W1s; float* const W1t = sm + S.W1t; float* const W2s = sm + S.W2s; float* const W2t = sm + S.W2t;
  float* const W3s = sm + S.W3s; float* const bs = sm + S.bs; float* const sF = sm + S.sF; float* const Fm = sm + S.F;
  float* const mF = sm + S.mF; float* const vF = sm + S.vF; float* const zw = sm + S.zs + warp * 128;
  float* const dE = sm + S.dE; float* const dZ3 = sm + S.dZ3; float* const logit = sm + S.logit;
  const bool wp_smem = C * (PD + 1) <= GX_WP_SMEM_MAX;
  const float* const Wpp = wp_smem ? sm + S.Wp : m.Wp;
  const float* const bpp = wp_smem ? sm + S.Wp + C * PD : m.bp;

  // model weights: once per CTA
  for (int idx = tid; idx < dp * HS; idx += NT) { const int f = idx / HS, c = idx - f * HS; W1s[idx] = f < d ? __ldg(m.W[0] + f * HID + c) : 0.f; }
  for (int idx = tid; idx < HID * dp; idx += NT) { const int c = idx / dp, f = idx - c * dp; W1t[idx] = f < d ? __ldg(m.Wt[0] + c * d + f) : 0.f; }
  for (int idx = tid; idx < HID * HS; idx += NT) { W2s[idx] = __ldg(m.W[1] + idx); W2t[idx] = __ldg(m.Wt[1] + idx); }
  for (int idx = tid; idx < HID * EMB; idx += NT) W3s[idx] = __ldg(m.W[2] + idx);
  for (int idx = tid; idx < HID; idx += NT) { bs[idx] = __ldg(m.b[0] + idx); bs[HID + idx] = __ldg(m.b[1] + idx); }
  for (int idx = tid; idx < EMB; idx += NT) bs[2 * HID + idx] = __ldg(m.b[2] + idx);
  if (wp_smem) {
    float* const Wps = sm + S.Wp;
    for (int idx = tid; idx < C * PD; idx += NT) Wps[idx] = __ldg(m.Wp + idx);
    for (int idx = tid; idx < C; idx += NT) Wps[C * PD + idx] = __ldg(m.bp + idx);
  }

  // lane groups of 8: a group owns one hid-wide row (q < H4 lanes carry a float4 each)
  Grp G;
  G.GW = 8; G.epi = 4; G.lane = lane; G.grp = lane >> 3; G.q = lane & 7; G.gbase = G.grp * 8;
  constexpr int epi = 4;
  const int q = G.q;
  float* const slab = A.gws + (int64_t)blockIdx.x * A.gws_stride_words;
  float2* const MM0 = reinterpret_cast<float2*>(A.pws + (int64_t)blockIdx.x * A.pws_stride_words);

  for (;;) {
    __syncthreads();
    if (tid == 0) s_task = atomicAdd(A.counter, 1);
    __syncthreads();
    const int qi = s_task;
    if (qi >= A.ntasks) break;
    const int task_id = A.order[qi];
    const GxTask* __restrict__ Tp = A.plan.tasks + task_id;
    const int n = Tp->n, n1 = Tp->n1, n2 = Tp->n2, e_d = Tp->e_d, np = Tp->npairs_in;
    const int gt = Tp->gt_label;
    const int64_t node_off = Tp->node_off, rp_off = Tp->rp_off, edge_off = Tp->edge_off, pair_off = Tp->pair_off;
    const GxStreamLayout L = gx_make_stream_layout(n, n1, n2, e_d, d, HID, nwarps);
    const int32_t* __restrict__ lo2gid = A.plan.lo2gid + node_off;
    const int32_t* const irp = A.plan.irowptr + rp_off;
    const int32_t* const icol = A.plan.icol + edge_off;
    const int32_t* __restrict__ pi = A.plan.pair_i + pair_off; const int32_t* __restrict__ pj = A.plan.pair_j + pair_off;
    const int32_t* __restrict__ ppij = A.plan.pair_pij + pair_off; const int32_t* __restrict__ ppji = A.plan.pair_pji + pair_off;
    const int32_t* __restrict__ poij = A.plan.pair_oij + pair_off; const int32_t* __restrict__ poji = A.plan.pair_oji + pair_off;
    float* const a = slab + L.a; float* const P = slab + L.P; float* const Yh1 = slab + L.Yh1; float* const q1 = slab + L.q1;
    float* const dY1 = slab + L.dY1; float* const Yh2 = slab + L.Yh2; float* const q2 = slab + L.q2; float* const dZ2 = slab + L.dZ2;
    float* const yv = slab + L.y; float* const gFp = slab + L.gFp;
    int32_t* const cnt1 = reinterpret_cast<int32_t*>(slab + L.cnt1); int32_t* const cnt2 = reinterpret_cast<int32_t*>(slab + L.cnt2);
    int32_t* const llist = reinterpret_cast<int32_t*>(slab + L.llist); int32_t* const llistB = reinterpret_cast<int32_t*>(slab + L.llistB);
    int32_t* const llistO = reinterpret_cast<int32_t*>(slab + L.llistO);
    float2* const MM = MM0; float2* const mm = MM + np; float2* const vv = mm + np; float2* const SS = vv + np;
    const float nn = (float)n * (float)n;
    const float ent_over_nn = hp.c_ent / nn;
    const float lap_over_nn = hp.c_lap / nn;

    // ------------------------------------------------------------------ per-task state
    for (int i = tid; i < n; i += NT) yv[i] = (float)__ldg(A.g.pred_label + lo2gid[i]);
    for (int f = tid; f < dp; f += NT) { sF[f] = 0.5f; Fm[f] = 0.f; mF[f] = 0.f; vF[f] = 0.f; }  // feat_mask = 0 (explain.py:633-643)
    {
      const float m0_std = sqrtf(2.0f / (float)n);  // gain('relu') * sqrt(2/(n+n)) (explain.py:647-651)
      for (int p = tid; p < np; p += NT) {
        const int oij = poij[p], oji = poji[p];
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
        const float a0 = 0.5f * (Si + Sj);  // explain.py:665-678
        a[ppij[p]] = a0;
        a[ppji[p]] = a0;
        if (hp.iters == 0) {
          A.out_mask[edge_off + oij] = a0;
          A.out_mask[edge_off + oji] = a0;
        }
      }
    }
    for (int i = tid; i < n; i += NT) {
      const int r0 = irp[i], r1 = irp[i + 1];
      cnt2[i] = prefix_below(icol, r0, r1, n2);
      if (i < n2) cnt1[i] = prefix_below(icol, r0, r1, n1);
    }
    __syncthreads();
    if (warp == 0) {
      int below;
      const int c = build_long_list(n2, n1, llist, lane, [&](int i) { return irp[i + 1] - irp[i]; }, &below);
      if (lane == 0) { s_long[0] = c; s_long[1] = below; }
    } else if (warp == 1) {
      int below;
      const int c = build_long_list(n2, 0, llistB, lane, [&](int i) { return cnt1[i]; }, &below);
      if (lane == 0) s_long[2] = c;
    } else if (warp == 2) {
      int below;
      const int c = build_long_list(n, 0, llistO, lane, [&](int i) { return cnt2[i]; }, &below);
      if (lane == 0) s_long[3] = c;
    }
    __syncthreads();
    const int nlongF1 = s_long[0], nlongF2 = s_long[1], nlongB1 = s_long[2], nlongO = s_long[3];

    // ------------------------------------------------------------------ epochs
    for (int it = 1; it <= hp.iters; ++it) {
      // ---- F0: all nodes: P = (X . sigmoid(feat_mask)) W1                         (explain.py:707, models.py:70-71)
      {
        const float4 s4 = lane < D4 ? ld4(sF + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int chunk = (D4 + epi - 1) / epi;            // the four lane groups split the feature axis
        const int f0 = G.grp * chunk, f1 = min(D4, f0 + chunk);
        for (int j = warp; j < n; j += nwarps) {
          if (lane < D4) {
            const float4 x = load_x4(A.g.feat + (int64_t)lo2gid[j] * d, lane, d, xvec);
            st4(zw + 4 * lane, make_float4(x.x * s4.x, x.y * s4.y, x.z * s4.z, x.w * s4.w));
          }
          __syncwarp();
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          if (q < H4 && f0 < f1) acc = group_dense(zw + 4 * f0, f1 - f0, W1s + 4 * f0 * HS, HS, q, acc);
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
          }
          if (G.grp == 0 && q < H4) st4(P + j * HS + 4 * q, acc);
          __syncwarp();
        }
      }
      __syncthreads();
      // ---- F1: rows [0,n2): Y1 = A_m P + b1 ; row normalise                                   (models.py:70-78)
      {
        const int ntask = nlongF1 + (n2 + epi - 1) / epi;
        for (int t = warp; t < ntask; t += nwarps) {
          float4 z;
          const int i = row_task_gather<int32_t, false, true>(t, nlongF1, llist, n2, G, H4, irp, icol, a, P, HS, (const int32_t*)nullptr, zw, z);
          const bool act = i >= 0;
          float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
          if (act && q < H4) { const float4 b = ld4(bs + 4 * q); y = make_float4(z.x + b.x, z.y + b.y, z.z + b.z, z.w + b.w); }
          const float ss = group_sum(y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w, G);
          const float qn = fmaxf(sqrtf(ss), 1e-12f);  // F.normalize(p=2, dim=2), eps 1e-12
          if (act && q < H4) st4(Yh1 + i * HS + 4 * q, make_float4(y.x / qn, y.y / qn, y.z / qn, y.w / qn));
          if (act && q == 0) q1[i] = qn;
        }
      }
      __syncthreads();
      // ---- F2: rows [0,n1): Y2 = (A_m relu(Yh1)) W2 + b2 ; row normalise
      {
        const int ntask = nlongF2 + (n1 + epi - 1) / epi;
        for (int t = warp; t < ntask; t += nwarps) {
          float4 z;
          const int i = row_task_gather<int32_t, true, true>(t, nlongF2, llist, n1, G, H4, irp, icol, a, Yh1, HS, (const int32_t*)nullptr, zw, z);
          const bool act = i >= 0;
          if (act && q < H4) st4(zw + lane * 4, z);
          __syncwarp();
          float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
          if (act && q < H4) y = group_dense(zw + G.gbase * 4, H4, W2s, HS, q, ld4(bs + HID + 4 * q));
          const float ss = group_sum(y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w, G);
          const float qn = fmaxf(sqrtf(ss), 1e-12f);
          if (act && q < H4) st4(Yh2 + i * HS + 4 * q, make_float4(y.x / qn, y.y / qn, y.z / qn, y.w / qn));
          if (act && q == 0) q2[i] = qn;
          __syncwarp();
        }
      }
      __syncthreads();
      // ---- S: row r (= level-order id 0): layer 3, readout, softmax, -log p[gt], layer-3 backward
      if (warp == 0) {
        {
          const int r0 = irp[0], r1 = irp[1];
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          if (q < H4) acc = gather_row<int32_t, true, true>(r0 + G.grp, r1, epi, icol, a, Yh2, HS, q);
          st4(zw + lane * 4, acc);
        }
        __syncwarp();
        float z = 0.f;
        if (lane < HID)
          for (int g2 = 0; g2 < epi; ++g2) z += zw[(g2 * 8 + (lane >> 2)) * 4 + (lane & 3)];
        __syncwarp();
        if (lane < HID) zw[lane] = z;
        __syncwarp();
        float y3 = lane < EMB ? bs[2 * HID + lane] : 0.f;
        if (lane < EMB)
          for (int f = 0; f < HID; ++f) y3 = fmaf(zw[f], W3s[f * EMB + lane], y3);
        const float ss = warp_sum(lane < EMB ? y3 * y3 : 0.f);
        const float q3 = fmaxf(sqrtf(ss), 1e-12f);
        const float yh3 = lane < EMB ? y3 / q3 : 0.f;
        const float e1v = lane < HID ? fmaxf(Yh1[lane], 0.f) : 0.f;  // row 0 of H1
        const float e2v = lane < HID ? fmaxf(Yh2[lane], 0.f) : 0.f;  // row 0 of H2
        // logits = pred_model(concat) (models.py:260,375), softmax over classes (explain.py:714)
        for (int c = 0; c < C; ++c) {
          const float* wp = Wpp + c * PD;
          float t = 0.f;
          if (lane < HID) t = fmaf(e1v, wp[lane], fmaf(e2v, wp[HID + lane], t));
          if (lane < EMB) t = fmaf(yh3, wp[2 * HID + lane], t);
          t = warp_sum(t);
          if (lane == 0) logit[c] = t + bpp[c];
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
          const float* wp = Wpp + c * PD;
          if (lane < HID) { d1 = fmaf(gc, wp[lane], d1); d2 = fmaf(gc, wp[HID + lane], d2); }
          if (lane < EMB) d3 = fmaf(gc, wp[2 * HID + lane], d3);
        }
        if (lane < HID) { dE[lane] = d1; dE[HS + lane] = d2; }
        // backward of y/max(|y|,eps): dY = (dYh - Yh <Yh,dYh>)/q ; dZ3 = dY3 W3^T
        const float s3 = warp_sum(yh3 * d3);
        const float dy3 = lane < EMB ? (d3 - yh3 * s3) / q3 : 0.f;
        __syncwarp();
        if (lane < EMB) zw[lane] = dy3;
        __syncwarp();
        if (lane < HID) dZ3[lane] = dot_v4(zw, W3s + lane * EMB, EMB / 4);
      }
      __syncthreads();
      // ---- B2: rows {r} U N(r): dYh2 = dEmb2 (row r) + a[r,j] dZ3 (j in N(r)), relu', normalise', dZ2 = dY2 W2^T
      {
        const int r0 = irp[0];
        const int items = 1 + irp[1] - r0;
        const int ntask = (items + epi - 1) / epi;
        for (int t = warp; t < ntask; t += nwarps) {
          const int item = t * epi + G.grp;
          const bool act = item < items;
          int j = 0;
          float coef = 1.f;
          const float* dsrc = dE + HS;
          if (act && item > 0) {
            const int e = r0 + item - 1;
            j = icol[e];
            coef = a[e];
            dsrc = dZ3;
          }
          float4 yh = make_float4(0.f, 0.f, 0.f, 0.f), dy = yh;
          if (act && q < H4) {
            yh = ld4(Yh2 + j * HS + 4 * q);
            const float4 g4 = ld4(dsrc + 4 * q);
            dy.x = yh.x > 0.f ? coef * g4.x : 0.f;   // relu backward: grad where input > 0
            dy.y = yh.y > 0.f ? coef * g4.y : 0.f;
            dy.z = yh.z > 0.f ? coef * g4.z : 0.f;
            dy.w = yh.w > 0.f ? coef * g4.w : 0.f;
          }
          const float sdot = group_sum(yh.x * dy.x + yh.y * dy.y + yh.z * dy.z + yh.w * dy.w, G);
          if (act && q < H4) {
            const float qn = q2[j];
            st4(zw + lane * 4, make_float4((dy.x - yh.x * sdot) / qn, (dy.y - yh.y * sdot) / qn,
                                           (dy.z - yh.z * sdot) / qn, (dy.w - yh.w * sdot) / qn));
          }
          __syncwarp();
          if (act && q < H4)
            st4(dZ2 + j * HS + 4 * q, group_dense(zw + G.gbase * 4, H4, W2t, HS, q, make_float4(0.f, 0.f, 0.f, 0.f)));
          __syncwarp();
        }
      }
      __syncthreads();
      // ---- B1: rows [0,n2): dH1 = A_m^T dZ2 (only columns < n1 carry gradient), relu', normalise' -> dY1
      {
        const int ntask = nlongB1 + (n2 + epi - 1) / epi;
        for (int t = warp; t < ntask; t += nwarps) {
          float4 dh;
          const int i = row_task_gather<int32_t, false, true>(t, nlongB1, llistB, n2, G, H4, irp, icol, a, dZ2, HS, cnt1, zw, dh);
          const bool act = i >= 0;
          float4 yh = make_float4(0.f, 0.f, 0.f, 0.f), dy = yh;
          if (act && q < H4) {
            yh = ld4(Yh1 + i * HS + 4 * q);
            if (i == 0) { const float4 e4 = ld4(dE + 4 * q); dh.x += e4.x; dh.y += e4.y; dh.z += e4.z; dh.w += e4.w; }
            dy.x = yh.x > 0.f ? dh.x : 0.f; dy.y = yh.y > 0.f ? dh.y : 0.f;
            dy.z = yh.z > 0.f ? dh.z : 0.f; dy.w = yh.w > 0.f ? dh.w : 0.f;
          }
          const float sdot = group_sum(yh.x * dy.x + yh.y * dy.y + yh.z * dy.z + yh.w * dy.w, G);
          if (act && q < H4) {
            const float qn = q1[i];
            st4(dY1 + i * HS + 4 * q, make_float4((dy.x - yh.x * sdot) / qn, (dy.y - yh.y * sdot) / qn,
                                                  (dy.z - yh.z * sdot) / qn, (dy.w - yh.w * sdot) / qn));
          }
        }
      }
      __syncthreads();
      // ---- B0: all nodes: dP = A_m^T dY1 (columns < n2 of row j), dL/dsF += X_j (.) (dP_j W1^T)
      {
        float4 gacc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int ntask = nlongO + (n + epi - 1) / epi;
        for (int t = warp; t < ntask; t += nwarps) {
          float4 z;
          const int i = row_task_gather<int32_t, false, true>(t, nlongO, llistO, n, G, H4, irp, icol, a, dY1, HS, cnt2, zw, z);
          for (int g2 = 0; g2 < epi; ++g2) {
            const int ig = __shfl_sync(0xffffffffu, i, g2 * 8);
            if (ig < 0) continue;  // warp-uniform
            if (G.grp == g2 && q < H4) st4(zw + 4 * q, z);
            __syncwarp();
            if (lane < D4) {
              const float4 o = group_dense(zw, H4, W1t, dp, lane, make_float4(0.f, 0.f, 0.f, 0.f));
              const float4 x = load_x4(A.g.feat + (int64_t)lo2gid[ig] * d, lane, d, xvec);
              gacc.x = fmaf(o.x, x.x, gacc.x); gacc.y = fmaf(o.y, x.y, gacc.y);
              gacc.z = fmaf(o.z, x.z, gacc.z); gacc.w = fmaf(o.w, x.w, gacc.w);
            }
            __syncwarp();
          }
        }
        if (lane < D4) st4(gFp + warp * dp + 4 * lane, gacc);  // per-warp partial, summed in warp order below
      }
      __syncthreads();
      // ---- P: per undirected edge: dA_ij, dA_ji, symmetrise, regularisers, Adam, next mask value
      {
        const float2 tab = __ldg(hp.adam_tab + (it - 1));
        const float step = tab.x, bc2s = tab.y, bc2s_inv = 1.0f / tab.y;
        const bool last = (it == hp.iters);
        // feature mask: dL/dF = sF(1-sF) (sum_j X_j (.) dX'_j + feat_size/d) ; Adam (explain.py:766, train_utils.py:10)
        for (int f = tid; f < d; f += NT) {
          float gsum = 0.f;
          for (int w = 0; w < nwarps; ++w) gsum += gFp[w * dp + f];
          const float s = sF[f];
          const float g = s * (1.f - s) * (gsum + hp.c_feat_size / (float)d);
          float mf = mF[f], vf = vF[f], Fv = Fm[f];
          mf = mf + (g - mf) * hp.one_minus_b1;
          vf = vf * hp.b2 + hp.one_minus_b2 * g * g;
          Fv = Fv - step * (mf / (sqrtf(vf) / bc2s + hp.eps));
          mF[f] = mf; vF[f] = vf; Fm[f] = Fv;
          sF[f] = sigmoid_f(Fv);
        }
        for (int p = tid; p < np; p += NT) {
          const int i = pi[p], j = pj[p];   // i < j, i < n2
          const float yd = yv[i] - yv[j];
          float Gd = lap_over_nn * yd * yd;  // d/dA_ij + d/dA_ji of y^T (D - A) y / n^2 (explain.py:780-793)
          Gd += dot_v4(dY1 + i * HS, P + j * HS, H4);
          if (j < n2) Gd += dot_v4(dY1 + j * HS, P + i * HS, H4);
          if (i < n1) Gd += dot_relu_v4(dZ2 + i * HS, Yh1 + j * HS, H4);
          if (j < n1) Gd += dot_relu_v4(dZ2 + j * HS, Yh1 + i * HS, H4);
          if (i == 0) Gd += dot_relu_v4(dZ3, Yh2 + j * HS, H4);
          Gd *= 0.5f;  // sym_mask = (S + S^T)/2 (explain.py:671)
          float2 Mv = MM[p];
          const float2 Sv = SS[p];
          // size: coeff*sum(S) ; entropy: mean over n^2 of H(S), dH/dM = -M S(1-S) (explain.py:755-770)
          const float gi = Sv.x * (1.f - Sv.x) * (Gd + hp.c_size - ent_over_nn * Mv.x);
          const float gj = Sv.y * (1.f - Sv.y) * (Gd + hp.c_size - ent_over_nn * Mv.y);
          float2 m2 = mm[p], v2 = vv[p];
          m2.x = m2.x + (gi - m2.x) * hp.one_minus_b1;
          m2.y = m2.y + (gj - m2.y) * hp.one_minus_b1;
          v2.x = v2.x * hp.b2 + hp.one_minus_b2 * gi * gi;
          v2.y = v2.y * hp.b2 + hp.one_minus_b2 * gj * gj;
          Mv.x = Mv.x - adam_delta_fast(m2.x, v2.x, step, bc2s_inv, hp.eps);
          Mv.y = Mv.y - adam_delta_fast(m2.y, v2.y, step, bc2s_inv, hp.eps);
          const float2 Sn = make_float2(sigmoid_fast(Mv.x), sigmoid_fast(Mv.y));
          MM[p] = Mv; mm[p] = m2; vv[p] = v2; SS[p] = Sn;
          const float an = 0.5f * (Sn.x + Sn.y);
          a[ppij[p]] = an;
          a[ppji[p]] = an;
          if (last) {
            A.out_mask[edge_off + poij[p]] = an;
            A.out_mask[edge_off + poji[p]] = an;
          }
        }
      }
      __syncthreads();
    }
    if (A.out_feat != nullptr)
      for (int f = tid; f < d; f += NT) A.out_feat[(int64_t)task_id * d + f] = sF[f];
  }
}

template <int HID, int EMB>
cudaError_t launch_stream(const GxExplainLaunch& cfg, const ExplainArgs& args, cudaStream_t s) {
  constexpr int NT = GX_STREAM_THREADS;
  auto kern = explain_stream_kernel<HID, EMB, NT>;
  const StreamSmem S = stream_smem(gx_round_up(args.m.d, 4), HID, EMB, args.m.C, NT / 32);
  const int bytes = S.total * 4;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  kern<<<cfg.grid, NT, bytes, s>>>(args);
  return cudaGetLastError();
}

}  // namespace

cudaError_t gx_launch_explain_stream(const GxExplainLaunch& cfg, const GxGraphDev& g, const GxModelDev& m,
                                     const GxHparamsDev& hp, const GxPlanArrays& plan, const float* m0,
                                     float* out_mask, float* out_feat, cudaStream_t s) {
  ExplainArgs args;
  args.order = cfg.order; args.ntasks = cfg.ntasks; args.counter = cfg.counter;
  args.gws = cfg.gws; args.gws_stride_words = cfg.gws_stride_words;
  args.pws = cfg.pws; args.pws_stride_words = cfg.pws_stride_words;
  args.g = g; args.m = m; args.hp = hp; args.plan = plan;
  args.m0 = m0; args.out_mask = out_mask; args.out_feat = out_feat; args.dbg = nullptr;
  if (m.hid == 20 && m.emb == 20) return launch_stream<20, 20>(cfg, args, s);
  if (m.hid == 32 && m.emb == 32) return launch_stream<32, 32>(cfg, args, s);   // any width <= 32, zero-padded by gx_set_model
  return cudaErrorInvalidValue;
}
