// forward.cu -- the GCN forward of the reference model on the WHOLE graph (models.py:58-80 GraphConv.forward, :230-267 gcn_forward,
// :363-376 GcnEncoderNode.forward): what produces the `pred` the Explainer is constructed with (explainer_main.py:186-193 reads it
// from the checkpoint; `Explainer(pred=None)` computes it here).  Unmasked adjacency, no feature mask, any num_layers <= 4, --bn.
// One launch per layer (a layer reads every row of the previous one), a warp per node with lane = feature -- the same row
// arithmetic as explain_var.cu: Y = (sum_{j in N(i)} H_{l-1}[j]) W_l + b_l, row L2-normalise, ReLU (+ per-node standardisation
// with --bn) on hidden layers; logits = pred_model(concat of the layer outputs).
#include <algorithm>

#include "explain_common.cuh"

namespace {

constexpr int kFwdThreads = 256;

// one GCN layer for all N nodes.  Hin: [N][32] (layer > 1) or the feature matrix [N][d] (layer 1); Hout: [N][32] what the next layer /
// the readout sees (relu / standardised on hidden layers, the normalised output on the last).
template <bool kFirst>
__global__ void __launch_bounds__(kFwdThreads) gcn_layer_kernel(GxGraphDev g, const float* __restrict__ Hin, const float* __restrict__ W, const float* __restrict__ b,
                                                                int win, int wout, int last, int bn, float* __restrict__ Hout) {
  extern __shared__ float zs_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kFwdThreads / 32;
  const int dp = gx_round_up(win, 4);
  float* const zs = zs_all + warp * dp;
  for (int64_t i = (int64_t)blockIdx.x * nwarps + warp; i < g.N; i += (int64_t)gridDim.x * nwarps) {
    const int r0 = g.rowptr[i], r1 = g.rowptr[i + 1];
    float y = lane < wout ? __ldg(b + lane) : 0.f;
    if (kFirst) {
      for (int f0 = 0; f0 < win; f0 += 32) {
        const int f = f0 + lane;
        float z = 0.f;
        if (f < win)
          for (int e = r0; e < r1; ++e) z += __ldg(Hin + (int64_t)g.col[e] * win + f);   // raw adjacency, self loops included (models.py:70: torch.matmul(adj, x))
        if (f < win) zs[f] = z;
      }
      __syncwarp();
      if (lane < wout)
        for (int f = 0; f < win; ++f) y = fmaf(zs[f], __ldg(W + f * wout + lane), y);
      __syncwarp();
    } else {
      float z = 0.f;
      if (lane < win)
        for (int e = r0; e < r1; ++e) z += Hin[(int64_t)g.col[e] * 32 + lane];
      for (int f = 0; f < win; ++f) {
        const float zf = __shfl_sync(0xffffffffu, z, f);
        if (lane < wout) y = fmaf(zf, __ldg(W + f * wout + lane), y);
      }
    }
    const float ss = warp_sum(lane < wout ? y * y : 0.f);
    const float q = fmaxf(sqrtf(ss), 1e-12f);   // F.normalize(p=2, dim=2), eps 1e-12 (models.py:78)
    float h = lane < wout ? y / q : 0.f;
    if (!last) {
      h = fmaxf(h, 0.f);
      if (bn) {   // fresh BatchNorm1d(n) in train mode: per node over the feature axis (models.py:222-228)
        const float mu = warp_sum(lane < wout ? h : 0.f) / (float)wout;
        const float dv = lane < wout ? h - mu : 0.f;
        const float var = warp_sum(dv * dv) / (float)wout;
        h = dv / sqrtf(var + 1e-5f);
      }
    }
    Hout[i * 32 + lane] = lane < wout ? h : 0.f;
  }
}

// logits[i][c] = bp[c] + sum_k emb_i[k] Wp[c][k], emb_i = [H_1[i] | ... | H_L[i]] (models.py:260,375)
__global__ void __launch_bounds__(kFwdThreads) readout_kernel(int64_t N, int L, int hid, int emb, int C, const float* __restrict__ H, const float* __restrict__ Wp,
                                                              const float* __restrict__ bp, float* __restrict__ pred, float* __restrict__ emb_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kFwdThreads / 32;
  const int PD = hid * (L - 1) + emb;
  for (int64_t i = (int64_t)blockIdx.x * nwarps + warp; i < N; i += (int64_t)gridDim.x * nwarps) {
    for (int c = 0; c < C; ++c) {
      float t = 0.f;
      for (int l = 0; l < L; ++l) {
        const int w = l == L - 1 ? emb : hid;
        if (lane < w) t = fmaf(H[((int64_t)l * N + i) * 32 + lane], __ldg(Wp + c * PD + hid * l + lane), t);
      }
      t = warp_sum(t);
      if (lane == 0) pred[i * C + c] = t + __ldg(bp + c);
    }
    if (emb_out != nullptr)
      for (int l = 0; l < L; ++l) {
        const int w = l == L - 1 ? emb : hid;
        if (lane < w) emb_out[i * PD + hid * l + lane] = H[((int64_t)l * N + i) * 32 + lane];
      }
  }
}

}  // namespace

// H: workspace [L][N][32] floats (device).  pred [N][C], emb_out [N][PD] or nullptr (device).
cudaError_t gx_launch_model_forward(const GxGraphDev& g, const GxModelDev& m, float* H, float* pred, float* emb_out, cudaStream_t s) {
  const int nwarps = kFwdThreads / 32;
  const int grid = (int)std::min<int64_t>((g.N + nwarps - 1) / nwarps, 148 * 8);
  for (int l = 0; l < m.L; ++l) {
    const int win = l == 0 ? m.d : m.hid, wout = l == m.L - 1 ? m.emb : m.hid;
    const float* Hin = l == 0 ? g.feat : H + (int64_t)(l - 1) * g.N * 32;
    float* Hout = H + (int64_t)l * g.N * 32;
    const size_t smem = (size_t)nwarps * gx_round_up(win, 4) * sizeof(float);
    if (l == 0) gcn_layer_kernel<true><<<grid, kFwdThreads, smem, s>>>(g, Hin, m.W[l], m.b[l], win, wout, l == m.L - 1, m.bn, Hout);
    else gcn_layer_kernel<false><<<grid, kFwdThreads, smem, s>>>(g, Hin, m.W[l], m.b[l], win, wout, l == m.L - 1, m.bn, Hout);
  }
  readout_kernel<<<grid, kFwdThreads, 0, s>>>(g.N, m.L, m.hid, m.emb, m.C, H, m.Wp, m.bp, pred, emb_out);
  return cudaGetLastError();
}
