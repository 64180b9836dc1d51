// explain_var.cu -- K2v: the mask-optimisation kernel for the model VARIANTS of the reference:
//   num_gc_layers = 2 / 3 / 4 (explainer_main.py:57-66, explain.py:64: n_hops = num_gc_layers; models.py:193-220,230-267) and
//   --bn (models.py:222-228: a FRESH BatchNorm1d(n) in train mode on the (1, n, h) activations = per-node standardisation over
//   the feature axis, eps 1e-5, biased variance, applied after the ReLU of every hidden layer; the readout concatenates the
//   standardised activations, models.py:241-260).
// It computes exactly what oracle/kernel_spec.py specifies (parameters on the directed edges, layer l only on the rows within
// L - l hops of the explained node, inner / outer pair split), for hidden / output widths up to 128 (the tuned kernels stop at 32) and d <= 128.  These options are rare, so the
// kernel is written for clarity, not speed: one persistent CTA per task, state in a per-CTA global slab (L2 resident for the
// reference's graph sizes), one warp per row with lane = feature, one thread per undirected edge in the edge phase.
// Phases per epoch (one __syncthreads each): F1 .. FL | S | BL .. B1 | P.  The default model (3 layers, no bn) never comes here.
#include "explain_common.cuh"

namespace {

constexpr int kVarThreads = 256;
constexpr int kVarWeightWords = 36 * 1024;   // conv weights are staged in shared memory up to this many floats (144 KB), read through L2 beyond

// KW = 32-lane chunks of a hidden-width row (lane = feature, chunk k holds features 32k + lane): 1 for widths <= 32, 2 <= 64, 4 <= 128
__host__ __device__ inline int var_kw(int hid, int emb) { const int w = hid > emb ? hid : emb; return w <= 32 ? 1 : (w <= 64 ? 2 : 4); }

struct VarSmem { int W[GX_MAX_LAYERS], b[GX_MAX_LAYERS], Wp, sF, F, mF, vF, zs, zlen, gFp, emb, dEmb, logit, w_in_smem, total; };
__host__ __device__ inline VarSmem var_smem(int d, int L, int hid, int emb, int C, int nwarps) {
  VarSmem S;
  const int dp = gx_round_up(d, 4);
  int o = 0;
  auto take = [&](int words) { int r = o; o += gx_round_up(words, 4); return r; };
  int wwords = 0;
  for (int l = 0; l < L; ++l) wwords += (l == 0 ? d : hid) * (l == L - 1 ? emb : hid);
  S.w_in_smem = wwords <= kVarWeightWords;
  for (int l = 0; l < L; ++l) {
    const int win = l == 0 ? d : hid, wout = l == L - 1 ? emb : hid;
    S.W[l] = take(S.w_in_smem ? win * wout : 0);
    S.b[l] = take(wout);
  }
  const int PD = hid * (L - 1) + emb;
  S.Wp = take(C * (PD + 1) <= GX_WP_SMEM_MAX ? C * (PD + 1) : 0);
  S.sF = take(dp); S.F = take(dp); S.mF = take(dp); S.vF = take(dp);
  S.zlen = dp > 32 * var_kw(hid, emb) ? dp : 32 * var_kw(hid, emb);   // per-warp scratch row: a feature row or a hidden row
  S.zs = take(nwarps * S.zlen);
  S.gFp = take(nwarps * dp);
  S.emb = take(PD); S.dEmb = take(PD); S.logit = take(C < 32 ? 32 : C);
  S.total = o;
  return S;
}

template <bool kBn, int KW>
__global__ void __launch_bounds__(kVarThreads) explain_var_kernel(const ExplainArgs A) {
  extern __shared__ __align__(16) float sm[];
  __shared__ int s_task;
  constexpr int NT = kVarThreads, nwarps = NT / 32;
  constexpr int VW = 32 * KW;   // row stride of every hidden-width array
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const GxModelDev& m = A.m;
  const GxHparamsDev& hp = A.hp;
  const int d = m.d, C = m.C, L = m.L, hid = m.hid, embw = m.emb;
  const int dp = gx_round_up(d, 4);
  const int PD = hid * (L - 1) + embw;
  const bool ieee = (hp.flags & GX_HP_IEEE_EDGE) != 0;
  const VarSmem S = var_smem(d, L, hid, embw, C, nwarps);
  float* const sF = sm + S.sF; float* const Fm = sm + S.F; float* const mF = sm + S.mF; float* const vF = sm + S.vF;
  float* const zs = sm + S.zs + warp * S.zlen;
  float* const gFp = sm + S.gFp;
  float* const emb = sm + S.emb; float* const dEmb = sm + S.dEmb; float* const logit = sm + S.logit;
  const bool wp_smem = C * (PD + 1) <= GX_WP_SMEM_MAX;
  const float* const Wpp = wp_smem ? sm + S.Wp : m.Wp;
  const float* const bpp = wp_smem ? sm + S.Wp + C * PD : m.bp;
  auto win_of = [&](int l) { return l == 0 ? d : hid; };            // l = 0 .. L-1
  auto wout_of = [&](int l) { return l == L - 1 ? embw : hid; };

  const float* Wl[GX_MAX_LAYERS];   // conv weights: shared memory when they fit, else global (L2 resident)
  for (int l = 0; l < L; ++l) {
    const int cnt = win_of(l) * wout_of(l);
    if (S.w_in_smem)
      for (int idx = tid; idx < cnt; idx += NT) sm[S.W[l] + idx] = __ldg(m.W[l] + idx);
    Wl[l] = S.w_in_smem ? sm + S.W[l] : m.W[l];
    for (int idx = tid; idx < wout_of(l); idx += NT) sm[S.b[l] + idx] = __ldg(m.b[l] + idx);
  }
  if (wp_smem) {
    for (int idx = tid; idx < C * PD; idx += NT) sm[S.Wp + idx] = __ldg(m.Wp + idx);
    for (int idx = tid; idx < C; idx += NT) sm[S.Wp + C * PD + idx] = __ldg(m.bp + idx);
  }
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
    const int n = Tp->n, n2 = Tp->n2, e1 = Tp->e1, np = Tp->npairs_in;
    const int gt = Tp->gt_label;
    const int64_t node_off = Tp->node_off, rp_off = Tp->rp_off, edge_off = Tp->edge_off, pair_off = Tp->pair_off;
    int R[GX_MAX_LAYERS + 1];   // R[l] = rows of layer l (1-based): nodes within L - l hops; R[0] = n
    R[0] = n;
    for (int l = 1; l <= L; ++l) R[l] = Tp->cum[L - l] < n ? Tp->cum[L - l] : n;
    const GxVarLayout Lo = gx_make_var_layout(n, n2, e1, np, d, L, VW);
    const int32_t* __restrict__ lo2gid = A.plan.lo2gid + node_off;
    const int32_t* __restrict__ irp = A.plan.irowptr + rp_off;
    const int32_t* __restrict__ icol = A.plan.icol + edge_off;
    const int32_t* __restrict__ pi = A.plan.pair_i + pair_off; const int32_t* __restrict__ pj = A.plan.pair_j + pair_off;
    const int32_t* __restrict__ ppij = A.plan.pair_pij + pair_off; const int32_t* __restrict__ ppji = A.plan.pair_pji + pair_off;
    const int32_t* __restrict__ poij = A.plan.pair_oij + pair_off; const int32_t* __restrict__ poji = A.plan.pair_oji + pair_off;
    float* const a = slab + Lo.a; float* const U = slab + Lo.U; float* const dZ1 = slab + Lo.dZ1; float* const lapg = slab + Lo.lapg;
    auto Yh = [&](int l) { return slab + Lo.Yh + (int64_t)(l - 1) * n2 * VW; };     // l = 1..L: normalised pre-activation
    auto Hh = [&](int l) { return slab + Lo.H + (int64_t)(l - 1) * n2 * VW; };      // l = 1..L: what the next layer / the readout sees
    auto dZ = [&](int l) { return slab + Lo.dZ + (int64_t)(l - 2) * n2 * VW; };     // l = 2..L: dL/d(A_m H_{l-1}) (width hid)
    auto qn = [&](int l) { return slab + Lo.q + (int64_t)(l - 1) * n2; };
    auto istd = [&](int l) { return slab + Lo.istd + (int64_t)(l - 1) * n2; };
    float2* const MM = MM0; float2* const mm = MM + np; float2* const vv = mm + np; float2* const SS = vv + np;
    const float nn = (float)n * (float)n;
    const float ent_over_nn = hp.c_ent / nn;
    const float lap_over_nn = hp.c_lap / nn;

    for (int f = tid; f < dp; f += NT) {
      sF[f] = 0.5f; Fm[f] = 0.f; mF[f] = 0.f; vF[f] = 0.f;   // feat_mask = 0 (explain.py:633-643)
      if (hp.out_iter == 0 && f < d && A.out_feat != nullptr) A.out_feat[(int64_t)task_id * d + f] = 0.5f;
    }
    {
      const float m0_std = sqrtf(2.0f / (float)n);  // gain('relu') * sqrt(2/(n+n)) (explain.py:647-651)
      for (int p = tid; p < np; p += NT) {
        const int oij = poij[p], oji = poji[p];
        float Mi, Mj;
        if (hp.init == GX_INIT_PHILOX) {
          Mi = 1.0f + m0_std * philox_normal(hp.seed, (uint32_t)Tp->node, (uint32_t)oij);
          Mj = 1.0f + m0_std * philox_normal(hp.seed, (uint32_t)Tp->node, (uint32_t)oji);
        } else {
          Mi = __ldg(A.m0 + edge_off + oij);
          Mj = __ldg(A.m0 + edge_off + oji);
        }
        MM[p] = make_float2(Mi, Mj);
        mm[p] = make_float2(0.f, 0.f);
        vv[p] = make_float2(0.f, 0.f);
        const float Si = sigmoid_f(Mi), Sj = sigmoid_f(Mj);
        SS[p] = make_float2(Si, Sj);
        const float a0 = 0.5f * (Si + Sj);  // explain.py:665-678
        const int i = pi[p], j = pj[p];
        if (i < n2) a[ppij[p]] = a0;
        if (j < n2) a[ppji[p]] = a0;
        const float yd = (float)__ldg(A.g.pred_label + lo2gid[i]) - (float)__ldg(A.g.pred_label + lo2gid[j]);
        lapg[p] = lap_over_nn * yd * yd;   // d/dA_ij + d/dA_ji of y^T (D - A) y / n^2 (explain.py:780-793)
        if (hp.out_iter == 0) { A.out_mask[edge_off + oij] = a0; A.out_mask[edge_off + oji] = a0; }
      }
    }
    __syncthreads();

    for (int it = 1; it <= hp.iters; ++it) {
      // ---------------------------------------------------------------- forward, layer by layer        (models.py:58-80,230-267)
      for (int l = 1; l <= L; ++l) {
        const int win = win_of(l - 1), wout = wout_of(l - 1);
        const float* const Ws = Wl[l - 1]; const float* const bsm = sm + S.b[l - 1];
        for (int i = warp; i < R[l]; i += nwarps) {
          const int r0 = irp[i], r1 = irp[i + 1];
          float y[KW];
#pragma unroll
          for (int k = 0; k < KW; ++k) y[k] = lane + 32 * k < wout ? bsm[lane + 32 * k] : 0.f;
          if (l == 1) {
            for (int f0 = 0; f0 < d; f0 += 32) {
              const int f = f0 + lane;
              float z = 0.f;
              if (f < d)
                for (int e = r0; e < r1; ++e) z = fmaf(a[e], __ldg(A.g.feat + (int64_t)lo2gid[icol[e]] * d + f), z);
              if (f < d) { U[(int64_t)i * dp + f] = z; zs[f] = z * sF[f]; }   // x * sigmoid(feat_mask) (explain.py:707), linear in x
            }
          } else {
            const float* const Hp = Hh(l - 1);
#pragma unroll
            for (int k = 0; k < KW; ++k) {
              const int f = lane + 32 * k;
              float z = 0.f;
              if (f < win)
                for (int e = r0; e < r1; ++e) z = fmaf(a[e], Hp[(int64_t)icol[e] * VW + f], z);
              if (f < win) zs[f] = z;
            }
          }
          __syncwarp();
          for (int f = 0; f < win; ++f) {
            const float zf = zs[f];
#pragma unroll
            for (int k = 0; k < KW; ++k)
              if (lane + 32 * k < wout) y[k] = fmaf(zf, Ws[f * wout + lane + 32 * k], y[k]);
          }
          __syncwarp();
          float ssl = 0.f;
#pragma unroll
          for (int k = 0; k < KW; ++k) ssl += lane + 32 * k < wout ? y[k] * y[k] : 0.f;
          const float ss = warp_sum(ssl);
          const float q = fmaxf(sqrtf(ss), 1e-12f);   // F.normalize(p=2, dim=2), eps 1e-12
          float yh[KW], h[KW];
#pragma unroll
          for (int k = 0; k < KW; ++k) { yh[k] = lane + 32 * k < wout ? y[k] / q : 0.f; h[k] = yh[k]; }
          if (l < L) {
#pragma unroll
            for (int k = 0; k < KW; ++k) h[k] = fmaxf(yh[k], 0.f);
            if (kBn) {   // fresh BatchNorm1d(n) in train mode: per node, over the feature axis (models.py:222-228)
              float sl = 0.f;
#pragma unroll
              for (int k = 0; k < KW; ++k) sl += lane + 32 * k < wout ? h[k] : 0.f;
              const float mu = warp_sum(sl) / (float)wout;
              float vl = 0.f;
#pragma unroll
              for (int k = 0; k < KW; ++k) { h[k] = lane + 32 * k < wout ? h[k] - mu : 0.f; vl += h[k] * h[k]; }
              const float var = warp_sum(vl) / (float)wout;
              const float is = 1.0f / sqrtf(var + 1e-5f);
#pragma unroll
              for (int k = 0; k < KW; ++k) h[k] *= is;
              if (lane == 0) istd(l)[i] = is;
            }
          }
#pragma unroll
          for (int k = 0; k < KW; ++k) {
            Yh(l)[(int64_t)i * VW + lane + 32 * k] = yh[k];
            Hh(l)[(int64_t)i * VW + lane + 32 * k] = lane + 32 * k < wout ? h[k] : 0.f;
          }
          if (lane == 0) qn(l)[i] = q;
        }
        __syncthreads();
      }
      // ---------------------------------------------------------------- S: readout of row r = level-order id 0, softmax, dEmb
      if (warp == 0) {
        for (int l = 1; l <= L; ++l)
          for (int c = lane; c < wout_of(l - 1); c += 32) emb[hid * (l - 1) + c] = Hh(l)[c];
        __syncwarp();
        for (int c = 0; c < C; ++c) {
          float t = 0.f;
          for (int k = lane; k < PD; k += 32) t = fmaf(emb[k], Wpp[c * PD + k], t);
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
        for (int c = lane; c < C; c += 32) logit[c] = expf(logit[c] - mx) / se - (c == gt ? 1.f : 0.f);  // p - onehot(gt) (explain.py:750-753)
        __syncwarp();
        for (int k = lane; k < PD; k += 32) {
          float t = 0.f;
          for (int c = 0; c < C; ++c) t = fmaf(logit[c], Wpp[c * PD + k], t);
          dEmb[k] = t;
        }
      }
      for (int idx = tid; idx < nwarps * dp; idx += NT) gFp[idx] = 0.f;
      __syncthreads();
      // ---------------------------------------------------------------- backward, layer by layer
      for (int l = L; l >= 1; --l) {
        const int win = win_of(l - 1), wout = wout_of(l - 1);
        const float* const Ws = Wl[l - 1];
        for (int i = warp; i < R[l]; i += nwarps) {
          // dL/dH_l[i] = (A_m^T dZ_{l+1})[i] over the neighbours that are rows of layer l+1 (a prefix of row i) + the readout's share
          float g[KW], yh[KW];
#pragma unroll
          for (int k = 0; k < KW; ++k) g[k] = 0.f;
          if (l < L) {
            const int r0 = irp[i], r1 = irp[i + 1], bound = R[l + 1];
            const float* const dZn = dZ(l + 1);
            for (int e = r0; e < r1; ++e) {
              const int j = icol[e];
              if (j >= bound) break;   // columns are partitioned by level
              const float ae = a[e];
#pragma unroll
              for (int k = 0; k < KW; ++k)
                if (lane + 32 * k < wout) g[k] = fmaf(ae, dZn[(int64_t)j * VW + lane + 32 * k], g[k]);
            }
          }
#pragma unroll
          for (int k = 0; k < KW; ++k) {
            if (i == 0 && lane + 32 * k < wout) g[k] += dEmb[hid * (l - 1) + lane + 32 * k];
            yh[k] = Yh(l)[(int64_t)i * VW + lane + 32 * k];
          }
          if (l < L) {
            if (kBn) {   // backward of the per-node standardisation: (g - mean(g) - Hb mean(g Hb)) * istd
              float hb[KW], s1 = 0.f, s2 = 0.f;
#pragma unroll
              for (int k = 0; k < KW; ++k) {
                hb[k] = Hh(l)[(int64_t)i * VW + lane + 32 * k];
                if (lane + 32 * k < wout) { s1 += g[k]; s2 += g[k] * hb[k]; }
              }
              const float m1 = warp_sum(s1) / (float)wout, m2 = warp_sum(s2) / (float)wout, is = istd(l)[i];
#pragma unroll
              for (int k = 0; k < KW; ++k) g[k] = lane + 32 * k < wout ? (g[k] - m1 - hb[k] * m2) * is : 0.f;
            }
#pragma unroll
            for (int k = 0; k < KW; ++k) g[k] = yh[k] > 0.f ? g[k] : 0.f;   // relu backward
          }
          float sl = 0.f;
#pragma unroll
          for (int k = 0; k < KW; ++k) sl += lane + 32 * k < wout ? yh[k] * g[k] : 0.f;
          const float sdot = warp_sum(sl);
          const float qi = qn(l)[i];
          __syncwarp();
#pragma unroll
          for (int k = 0; k < KW; ++k)
            if (lane + 32 * k < wout) zs[lane + 32 * k] = (g[k] - yh[k] * sdot) / qi;   // dY: backward of y / max(|y|, eps)
          __syncwarp();
          // dZ[f] = sum_c dY[c] W[f][c]
          if (l == 1) {
            for (int f0 = 0; f0 < d; f0 += 32) {
              const int f = f0 + lane;
              float t = 0.f;
              if (f < d)
                for (int c = 0; c < wout; ++c) t = fmaf(zs[c], Ws[f * wout + c], t);
              if (f < d) {
                gFp[warp * dp + f] = fmaf(t, U[(int64_t)i * dp + f], gFp[warp * dp + f]);   // dL/dsF partial (U = A_m X)
                dZ1[(int64_t)i * dp + f] = t * sF[f];                                        // kept masked for the edge dots
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < KW; ++k) {
              const int f = lane + 32 * k;
              float t = 0.f;
              if (f < win)
                for (int c = 0; c < wout; ++c) t = fmaf(zs[c], Ws[f * wout + c], t);
              dZ(l)[(int64_t)i * VW + f] = f < win ? t : 0.f;
            }
          }
          __syncwarp();
        }
        __syncthreads();
      }
      // ---------------------------------------------------------------- P: edge gradients, regularisers, Adam, next mask
      {
        const float2 tab = __ldg(hp.adam_tab + (it - 1));
        const float step = tab.x, bc2s = tab.y, bc2s_inv = 1.0f / tab.y;
        const bool last = (it == hp.out_iter);
        for (int f = tid; f < d; f += NT) {
          float gsum = 0.f;
          for (int w = 0; w < nwarps; ++w) gsum += gFp[w * dp + f];
          const float s = sF[f];
          const float g = s * (1.f - s) * (gsum + hp.c_feat_size / (float)d);
          float mf = mF[f], vf = vF[f], Fv = Fm[f];
          if (hp.opt == GX_OPT_ADAM) {
            mf = mf + (g - mf) * hp.one_minus_b1;
            vf = vf * hp.b2 + hp.one_minus_b2 * g * g;
            Fv = Fv - step * (mf / (sqrtf(vf) / bc2s + hp.eps));
          } else {
            opt_step_other(hp.opt, Fv, g, mf, vf, step);
          }
          mF[f] = mf; vF[f] = vf; Fm[f] = Fv;
          const float sn = sigmoid_f(Fv);
          sF[f] = sn;   // (the edge dots below use dZ1 (.) sF stored in the backward, not this value)
          if (last && A.out_feat != nullptr) A.out_feat[(int64_t)task_id * d + f] = sn;
        }
        for (int p = tid; p < np; p += NT) {
          const int i = pi[p], j = pj[p];   // i < j in level order
          float Gd = lapg[p];
          if (i < R[1]) {
            float t = 0.f;
            const float* xr = A.g.feat + (int64_t)lo2gid[j] * d;
            for (int f = 0; f < d; ++f) t = fmaf(dZ1[(int64_t)i * dp + f], __ldg(xr + f), t);
            Gd += t;
          }
          if (j < R[1]) {
            float t = 0.f;
            const float* xr = A.g.feat + (int64_t)lo2gid[i] * d;
            for (int f = 0; f < d; ++f) t = fmaf(dZ1[(int64_t)j * dp + f], __ldg(xr + f), t);
            Gd += t;
          }
          for (int l = 2; l <= L; ++l) {
            const float* const dZl = dZ(l); const float* const Hp = Hh(l - 1);
            if (i < R[l]) { float t = 0.f; for (int f = 0; f < hid; ++f) t = fmaf(dZl[(int64_t)i * VW + f], Hp[(int64_t)j * VW + f], t); Gd += t; }
            if (j < R[l]) { float t = 0.f; for (int f = 0; f < hid; ++f) t = fmaf(dZl[(int64_t)j * VW + f], Hp[(int64_t)i * VW + f], t); Gd += t; }
          }
          Gd *= 0.5f;  // sym_mask = (S + S^T)/2 (explain.py:671)
          float2 Mv = MM[p];
          const float2 Sv = SS[p];
          float2 m2 = mm[p], v2 = vv[p];
          const float gi = Sv.x * (1.f - Sv.x) * (Gd + hp.c_size - ent_over_nn * Mv.x);
          const float gj = Sv.y * (1.f - Sv.y) * (Gd + hp.c_size - ent_over_nn * Mv.y);
          if (hp.opt == GX_OPT_ADAM) {
            m2.x = m2.x + (gi - m2.x) * hp.one_minus_b1;
            m2.y = m2.y + (gj - m2.y) * hp.one_minus_b1;
            v2.x = v2.x * hp.b2 + hp.one_minus_b2 * gi * gi;
            v2.y = v2.y * hp.b2 + hp.one_minus_b2 * gj * gj;
            Mv.x = Mv.x - adam_delta_fast(m2.x, v2.x, step, bc2s, bc2s_inv, hp.eps, ieee);
            Mv.y = Mv.y - adam_delta_fast(m2.y, v2.y, step, bc2s, bc2s_inv, hp.eps, ieee);
          } else {
            opt_step_other(hp.opt, Mv.x, gi, m2.x, v2.x, step);
            opt_step_other(hp.opt, Mv.y, gj, m2.y, v2.y, step);
          }
          const float2 Sn = make_float2(sigmoid_fast(Mv.x, ieee), sigmoid_fast(Mv.y, ieee));
          MM[p] = Mv; mm[p] = m2; vv[p] = v2; SS[p] = Sn;
          const float an = 0.5f * (Sn.x + Sn.y);
          if (i < n2) a[ppij[p]] = an;
          if (j < n2) a[ppji[p]] = an;
          if (last) { A.out_mask[edge_off + poij[p]] = an; A.out_mask[edge_off + poji[p]] = an; }
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace

int gx_var_smem_bytes(int d, int L, int hid, int emb, int C) { return var_smem(d, L, hid, emb, C, kVarThreads / 32).total * 4; }
int gx_var_row_stride(int hid, int emb) { return 32 * var_kw(hid, emb); }

cudaError_t gx_launch_explain_var(const GxExplainLaunch& cfg, const GxGraphDev& g, const GxModelDev& m,
                                  const GxHparamsDev& hp, const GxPlanArrays& plan, const float* m0,
                                  float* out_mask, float* out_feat, cudaStream_t s) {
  ExplainArgs args;
  args.order = cfg.order; args.ntasks = cfg.ntasks; args.counter = cfg.counter;
  args.gws = cfg.gws; args.gws_stride_words = cfg.gws_stride_words;
  args.pws = cfg.pws; args.pws_stride_words = cfg.pws_stride_words;
  args.g = g; args.m = m; args.hp = hp; args.plan = plan;
  args.m0 = m0; args.out_mask = out_mask; args.out_feat = out_feat; args.dbg = cfg.dbg; args.x = cfg.x;
  const int bytes = gx_var_smem_bytes(m.d, m.L, m.hid, m.emb, m.C);
  const int kw = var_kw(m.hid, m.emb);
  auto go = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    // every launch class asks for the largest shared-memory carveout: CTAs of different classes (= different kernels / footprints) can then
    // share an SM; with per-kernel carveouts a CTA waits for an SM that is completely idle (profiles/r02cl_cluster_auto.md)
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    kern<<<cfg.grid, kVarThreads, bytes, s>>>(args);
    return cudaGetLastError();
  };
  if (m.bn) {
    if (kw == 1) return go(explain_var_kernel<true, 1>);
    if (kw == 2) return go(explain_var_kernel<true, 2>);
    return go(explain_var_kernel<true, 4>);
  }
  if (kw == 1) return go(explain_var_kernel<false, 1>);
  if (kw == 2) return go(explain_var_kernel<false, 2>);
  return go(explain_var_kernel<false, 4>);
}
