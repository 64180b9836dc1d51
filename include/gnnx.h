/* gnnx.h -- C ABI of libgnnx.so, the B200-native GNNExplainer mask-optimisation engine.
 *
 * The reference (RexYing/gnn-model-explainer) has NO FFI/plugin interface: its boundary for this
 * hot path is a Python surface.  Every entry point below therefore cites the reference Python
 * function it replaces; INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - C linkage, no exceptions cross the ABI.  Every call returns GX_OK (0) or a negative
 *     gx_status; gx_last_error() returns a thread-local human-readable message.
 *   - The caller owns every input/output buffer.  The library owns only the opaque gx_handle
 *     (device copies of model/graph, the extraction plan and its workspace).
 *   - Pointers are HOST pointers unless the parameter is documented "device" or the call takes a
 *     gx_memspace.  No torch types appear in any signature.
 *   - One handle per host thread / per GPU.  Calls on one handle must not overlap.
 *   - All work is issued on the stream set with gx_set_stream (default: the legacy default
 *     stream); calls that return results to host memory synchronise that stream before returning.
 *   - There is NO CPU fallback: without a CUDA device gx_create fails with GX_ERR_CUDA.
 */
#ifndef GNNX_H_
#define GNNX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GX_VERSION 100

typedef struct gx_handle gx_handle;

typedef enum gx_status {
  GX_OK = 0,
  GX_ERR_INVALID = -1,     /* bad argument / call order                         */
  GX_ERR_CUDA = -2,        /* CUDA runtime error (message has the cudaError)    */
  GX_ERR_UNSUPPORTED = -3, /* valid in the reference but not built here (yet)   */
  GX_ERR_NODE = -4,        /* a node is not inside its own k-hop neighbourhood  */
  GX_ERR_NOMEM = -5
} gx_status;

typedef enum gx_memspace { GX_HOST = 0, GX_DEVICE = 1 } gx_memspace;

/* Model dimensions: GcnEncoderNode/GcnEncoderGraph(input_dim, hidden_dim, embedding_dim, label_dim,
 * num_layers, bn=..., args.bias) -- reference models.py:84-97,332-345. */
typedef struct gx_model_dims {
  int32_t input_dim;   /* d                                             */
  int32_t hidden_dim;  /* output width of conv_first / conv_block[*]    */
  int32_t embed_dim;   /* output width of conv_last                     */
  int32_t num_classes; /* label_dim                                     */
  int32_t num_layers;  /* num_gc_layers; this build: 3                  */
  int32_t flags;       /* GX_MODEL_* bits                               */
} gx_model_dims;
#define GX_MODEL_BN 1u /* args.bn (models.py:222-228): GX_ERR_UNSUPPORTED in this build */

/* Optimisation hyper-parameters: explainer_main.py:143-167 defaults + ExplainModule.coeffs
 * (explainer/explain.py:624-631) + torch.optim.Adam defaults (utils/train_utils.py:10). */
typedef struct gx_hparams {
  int32_t num_epochs;    /* args.num_epochs, default 100                                  */
  float lr;              /* args.lr, default 0.1                                          */
  float beta1, beta2;    /* 0.9, 0.999                                                    */
  float eps;             /* 1e-8                                                          */
  float coef_size;       /* coeffs["size"] = 0.005                                        */
  float coef_feat_size;  /* coeffs["feat_size"] = 1.0                                     */
  float coef_ent;        /* coeffs["ent"] = 1.0                                           */
  float coef_lap;        /* coeffs["lap"] = 1.0 (forced to 0 in graph mode)               */
  int32_t mask_act;      /* 0 = sigmoid (args.mask_act default); others GX_ERR_UNSUPPORTED */
  int32_t mask_bias;     /* args.mask_bias: accepted; a no-op exactly as in the reference (bias stays 0: ReLU6'(0)=0) */
  int32_t init;          /* GX_INIT_*                                                     */
  uint64_t seed;         /* GX_INIT_PHILOX: stream seed                                   */
} gx_hparams;
#define GX_INIT_M0 0     /* caller supplies M0 at the directed-edge entries (parity with torch's RNG draw) */
#define GX_INIT_PHILOX 1 /* N(1, 2/n) drawn on device, counter = (seed, node, edge slot)                  */

void gx_default_hparams(gx_hparams* hp);

const char* gx_last_error(void);
int gx_version(void);

/* Lifetime.  device = CUDA ordinal. */
int gx_create(int device, gx_handle** out);
int gx_destroy(gx_handle* h);
int gx_set_stream(gx_handle* h, void* cuda_stream);
int gx_sync(gx_handle* h);

/* Frozen model being explained.  Replaces the torch module the reference passes to
 * Explainer(model=...) (explain.py:43-57); tensors are the state_dict entries
 * conv_first.weight (d,h) / conv_block.0.weight (h,h) / conv_last.weight (h,e) as row-major
 * (in,out) float32, their biases (NULL = --nobias), pred_model.weight (C, 2h+e) row-major and
 * pred_model.bias (C).  conv_w / conv_b are arrays of num_layers pointers. */
int gx_set_model(gx_handle* h, const gx_model_dims* dims, const float* const* conv_w,
                 const float* const* conv_b, const float* pred_w, const float* pred_b);

/* Graph of a node-classification task, replacing Explainer(adj, feat, label, pred) (explain.py:43-62):
 * CSR of the (B=1) adjacency with ascending columns per row (must be symmetric 0/1; self loops are
 * honoured by gx_plan_nodes' reachability and dropped from the explained edge set exactly like the
 * reference's diag_mask, explain.py:617,678), features (N,d) float32, label (N) and
 * pred_label = argmax(pred[0], axis=1) (N) (explain.py:105). */
int gx_set_graph_csr(gx_handle* h, int64_t num_nodes, const int32_t* rowptr, const int32_t* col,
                     const float* feat, int32_t feat_dim, const int32_t* label,
                     const int32_t* pred_label);

/* graph_utils.neighborhoods (utils/graph_utils.py:147-158) for a set of rows: writes row `nodes[t]`
 * of the dense 0/1 hop matrix into out[t*num_nodes .. ] (uint8).  Integer BFS on CSR; bit-exact. */
int gx_neighborhood_rows(gx_handle* h, const int32_t* nodes, int32_t count, int32_t n_hops,
                         uint8_t* out_rows);

/* Explainer.extract_neighborhood (explain.py:492-501) for a batch of nodes, on device.
 * Builds the extraction plan kept inside the handle and reports the packed sizes:
 *   total_nodes = sum_t n_t,  total_edges = sum_t E_t (directed entries of the induced sub-adjacency).
 * Fails with GX_ERR_NODE if some node is outside its own neighbourhood (isolated node / n_hops=1
 * without self loop: the reference then explains a wrong row or crashes, explain.py:496-501). */
int gx_plan_nodes(gx_handle* h, const int32_t* nodes, int32_t count, int32_t n_hops,
                  int64_t* total_nodes, int64_t* total_edges);

/* Copies the canonical (reference-ordered) description of the planned subgraphs to the host:
 *   node_off[count+1], edge_off[count+1]      packed offsets
 *   neighbors[total_nodes]                    ascending global ids           (explain.py:497)
 *   node_idx_new[count]                       rank of the node in its set    (explain.py:496)
 *   sub_rowptr[total_nodes+count]             per task n_t+1 entries, task-local, starting at 0
 *   sub_col[total_edges]                      local column ids, ascending per row: the row-major
 *                                             nonzero order of the reference's dense sub_adj
 * Any pointer may be NULL to skip that array. */
int gx_plan_fetch(gx_handle* h, int64_t* node_off, int64_t* edge_off, int32_t* neighbors,
                  int32_t* node_idx_new, int32_t* sub_rowptr, int32_t* sub_col);

/* The hot path: Explainer.explain's optimisation loop for every planned node
 * (explain.py:97-146,209-211; ExplainModule explain.py:583-808; models.py:58-80,230-267,363-376;
 * torch.optim.Adam), one persistent CTA per node, all epochs in one launch.
 *   m0_edges    [total_edges] float32 in `space`: M0[i,j] at the sub_col slots (GX_INIT_M0), or NULL
 *   edge_mask   [total_edges] float32 in `space`: returned masked_adj[i,j] at the same slots
 *   feat_mask   optional [count*input_dim] float32 in `space`: sigmoid(feat_mask) after the last
 *               observed update (not returned by the reference API; for tests), may be NULL */
int gx_explain_nodes(gx_handle* h, const gx_hparams* hp, gx_memspace space, const float* m0_edges,
                     float* edge_mask, float* feat_mask);

/* The gradient baseline, Explainer.explain(..., model="grad") (explain.py:125-133) with ExplainModule.adj_feat_grad
 * (explain.py:717-738), for every planned node: one forward of the frozen model on the unmasked sub-adjacency and
 * features, loss = -log softmax(logits[node])[predicted label of the node], one backward to the adjacency;
 *   edge_mask [total_edges] float32 in `space`: sigmoid(|dL/dA_ij| + |dL/dA_ji|) at the sub_col slots. */
int gx_grad_nodes(gx_handle* h, gx_memspace space, float* edge_mask);

/* ---- graph-classification mode (Explainer(..., graph_mode=True), explain_graphs: explain.py:80-85,356-363) ----
 * Batch of padded graphs, replacing Explainer(adj (G,n,n), feat (G,n,d), label (G)): block CSR over
 * G*max_nodes rows (rowptr[G*max_nodes+1] with global edge offsets, col = node id inside its graph,
 * ascending per row, symmetric 0/1, no self loops), features (G*max_nodes, d), one label per graph. */
int gx_set_graph_batch_csr(gx_handle* h, int32_t num_graphs, int32_t max_nodes, const int32_t* rowptr,
                           const int32_t* col, const float* feat, int32_t feat_dim, const int32_t* label);
/* Plans the graphs to explain; edge_off[count+1] (may be NULL) receives the packed slot offsets: the slots of
 * graph t are the entries of its adjacency in row-major order (its slice of the CSR). */
int gx_plan_graphs(gx_handle* h, const int32_t* graph_ids, int32_t count, int64_t* edge_off, int64_t* total_edges);
/* Explainer.explain(node_idx=0, graph_idx=g, graph_mode=True) for every planned graph (model =
 * GcnEncoderGraph: per-layer max-pool readout, models.py:269-316; lap_loss = 0, explain.py:787-788).
 * m0_edges / edge_mask: [total_edges] in `space`, as for gx_explain_nodes. */
int gx_explain_graphs(gx_handle* h, const gx_hparams* hp, gx_memspace space, const float* m0_edges,
                      float* edge_mask, float* feat_mask);

/* Expands packed edge masks to the dense (n_t, n_t) float64 arrays Explainer.explain returns
 * (explain.py:209-221), task after task, into out (sum_t n_t^2 doubles, `space`). */
int gx_densify(gx_handle* h, gx_memspace space, const float* edge_mask, double* out);

/* Counters for bench.py: number of kernels this handle has launched so far, and the device time
 * (CUDA events on the handle's streams) of the explainer kernels of the last gx_explain_nodes call. */
int64_t gx_launch_count(gx_handle* h);
int gx_last_explain_ms(gx_handle* h, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* GNNX_H_ */
