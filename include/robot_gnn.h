/*
 * robot_gnn.h - C ABI of librobotgnn.so: B200 (sm_100a) kernels for the robot-graph
 * message-passing hot path of xjh19971/multi-robot-perception-gnn-1.
 *
 * Conventions:
 *   - every pointer is a CALLER-OWNED DEVICE pointer (torch tensors' data_ptr()); the
 *     library never allocates or frees user-visible memory - scratch is passed in;
 *   - every call takes sizes and a `stream` (cudaStream_t cast to void*), enqueues
 *     work on it and returns without a host sync;
 *   - return value: 0 = ok, <0 = invalid argument (RGNN_E_*), >0 = cudaError_t;
 *     rgnn_last_error() gives a thread-local message for the last non-zero return;
 *   - dtype codes: RGNN_F32 = 0 (parity mode), RGNN_BF16 = 1 (throughput mode; fp32
 *     accumulation everywhere); integer graph arrays are int64 where DGL exposes them
 *     (bit-exact contract) and int32 in the compact copies the compute kernels walk;
 *   - feature maps are NCHW flattened to [N, F, P] (P = h*w, contiguous), exactly the
 *     memory the reference hands to the layer (dgl/model/models.py:179-180);
 *   - no CPU fallback exists: with no CUDA device every compute entry returns an error.
 *
 * Each entry cites the reference interface (file:line under /root/reference) it replaces.
 */
#ifndef ROBOT_GNN_H_
#define ROBOT_GNN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGNN_VERSION 100

#define RGNN_F32 0
#define RGNN_BF16 1

#define RGNN_E_ARG (-1)       /* null pointer / negative size / bad enum */
#define RGNN_E_ALIGN (-2)     /* pointer or row size not aligned as the kernel requires */
#define RGNN_E_RANGE (-3)     /* size exceeds what the kernel supports (e.g. >= 2^31 edges) */
#define RGNN_E_UNSUPPORTED (-4)

/* reduce modes of rgnn_aggregate_* */
#define RGNN_POST_NONE 0      /* plain sum                                   (fn.sum)            */
#define RGNN_POST_MEAN 1      /* sum / in-degree, 0 for isolated nodes       (mailbox.mean(1))   */
#define RGNN_POST_SCALE 2     /* sum * dst_scale[v]                          (GraphConv norm)    */

/* epilogue flags of rgnn_gemm (bit-or):  v = alpha*acc [+ beta*aux[m,n]] [+ bias_m[m]] [+ bias_n[n]];
 * then optionally sigmoid(v), or v * (aux[m,n] > 0) */
#define RGNN_EPI_BIAS_M 1     /* + bias[m]            (1x1 conv bias, models.py:164)              */
#define RGNN_EPI_BIAS_N 2     /* + bias[n]            (nn.Linear bias, models.py:149)             */
#define RGNN_EPI_SIGMOID 4    /* sigmoid(v)           (edge encoder, models.py:150)               */
#define RGNN_EPI_RELU_MASK 8  /* v * (aux[m,n] > 0)   (ReLU backward, models.py:148)              */
#define RGNN_EPI_ADD_AUX 16   /* + beta*aux[m,n]      (accumulate / GCN2Conv identity mapping)    */

/* ---- library queries ------------------------------------------------------------------- */
int rgnn_version(void);                 /* RGNN_VERSION */
int rgnn_arch(void);                    /* 100: compiled for sm_100a only */
const char* rgnn_last_error(void);      /* thread-local, never NULL */
int rgnn_device_ok(void);               /* 0 if the current device is sm_100; else error code */
long long rgnn_launch_count(void);      /* kernels this library has launched in this process (bench evidence) */

/* ---- A2: dgl.batch index work  (dgl/training.py:57-58, dgl/eval.py:54-55) ---------------
 * local_src/local_dst [E]: per-graph LOCAL node ids concatenated in graph order.
 * batch_num_nodes/batch_num_edges [G].  Outputs: node_offsets/edge_offsets [G+1]
 * (exclusive scans), src/dst [E] = local id + node offset of the owning graph. */
int rgnn_batch_graphs(const int64_t* local_src, const int64_t* local_dst,
                      const int64_t* batch_num_nodes, const int64_t* batch_num_edges,
                      int64_t num_graphs, int64_t num_edges,
                      int64_t* node_offsets, int64_t* edge_offsets,
                      int64_t* src, int64_t* dst, void* stream);

/* ---- A3: in-degree / in-edge lookup / GraphConv norm -----------------------------------
 * Replaces what DGL does inside update_all (call site dgl/model/models.py:223): in_degrees,
 * in_edges(form='eid') sorted ascending, and GraphConv's degree norms.
 * src/dst [E] global ids.  node_offsets/edge_offsets [G+1] describe the batch (every edge of
 * graph g joins two nodes of graph g); pass G = 1 with offsets {0,N},{0,E} for an unbatched graph.
 * Outputs (int64, DGL-visible, bit-exact): in_deg/out_deg [N]; in_ptr/out_ptr [N+1];
 * in_src/in_eid, out_dst/out_eid [E] (stable: ascending edge id inside a node).
 * Outputs (int32 compact copies walked by the kernels), packed back to back:
 *   ptr32 [2(N+1) + G+1] = {in_ptr32 [N+1], out_ptr32 [N+1], node_offsets32 [G+1]},
 *   idx32 [6E] = {in_src32, in_eid32, out_dst32, out_eid32, src32, dst32} each [E].
 * norm [3N] fp32 = {1/in_deg (0 if isolated), clamp(in_deg,1)^-1/2, clamp(out_deg,1)^-1/2}.
 * status [1] int32: set non-zero on device if an edge leaves its graph's node range.
 * Any int64 output may be NULL (skipped). */
int rgnn_build_batched_csr(const int64_t* src, const int64_t* dst, int64_t num_nodes, int64_t num_edges,
                           const int64_t* node_offsets, const int64_t* edge_offsets, int64_t num_graphs,
                           int64_t* in_deg, int64_t* out_deg,
                           int64_t* in_ptr, int64_t* in_src, int64_t* in_eid,
                           int64_t* out_ptr, int64_t* out_dst, int64_t* out_eid,
                           int32_t* ptr32, int32_t* idx32, float* norm, int32_t* status, void* stream);

/* ---- A1: topology check -------------------------------------------------------------------------------------
 * ORs 2 into *status (device int32) unless src/dst (global ids, [E]) are exactly num_graphs complete directed
 * `robots`-node graphs without self loops in the reference's edge order: eid(i->j) = i*(K-1) + (j<i ? j : j-1)
 * (dgl/dataloader.py:88-95).  Validates the complete_k fast path of rgnn_aggregate / rgnn_film_backward. */
int rgnn_check_complete(const int64_t* src, const int64_t* dst, int64_t num_graphs, int64_t robots, int64_t num_edges,
                        int32_t* status, void* stream);

/* ---- next-row 3: edge features from raw robot poses (dgl/utils.py:54-77, dgl/dataloader.py:116-122) -------------
 * poses [G, K, 7] (xyz + quaternion xyzw; fp64 if pose_is_f64 else fp32) -> edge_pose [G*K*(K-1), 9] fp32 in the reference's
 * edge order: (t_j - t_i) || first six column-major entries of R(q_j * conj(q_i)).  fp64 arithmetic in the reference's
 * operation order, rounded to fp32 at the end: bit-identical to cal_relative_pose(...).float(). */
int rgnn_relative_pose(int pose_is_f64, const void* poses, float* edge_pose, int64_t num_graphs, int64_t robots, void* stream);

/* ---- A5/A6 (+A10, A11): fused gather -> FiLM / normalise -> sum|mean ----------------------
 * out[v,c,p] = post_v( sum_{i in [ptr[v],ptr[v+1])} ( gamma(eid_i,c) * nbr_scale[nbr_i] * x[nbr_i,c,p]
 *                                                    + beta(eid_i,c) ) ) + addend[v,c,p]
 * Replaces edge_udf + node_udf + degree bucketing (dgl/model/models.py:207-211,223) - the
 * [E,F,P] messages are never materialised - and update_all(copy_u, sum|mean)
 * (dgl/model/dgl_models.py:126; GraphConv/GCN2Conv).  Summation order is ascending edge id and
 * gamma*x, +beta are two roundings, so RGNN_F32 + RGNN_POST_MEAN reproduces the reference bit for bit.
 *   x [N, F, P] with node stride ldx elements;  z [E, 2F] interleaved (gamma,beta) or NULL (gamma=1,beta=0);
 *   use_beta = 0 ignores beta (the transposed pass of the backward);  nbr_scale [N] fp32 or NULL;
 *   dst_scale [N] fp32 (RGNN_POST_SCALE);  addend [N,F,P] node stride ldadd, or NULL;
 *   out node stride ldout (lets the result land in the 2nd half of a concat buffer);
 *   copy_x_to: if non-NULL also writes x[v] to copy_x_to[v] (node stride ldout) = torch.cat((h,g_h),1)
 *   (dgl/model/models.py:182) for free.
 *   ptr32 [N+1], nbr32 [E], eid32 [E]: the in-CSR (forward) or out-CSR (backward).
 *   node_offsets32 [G+1] or NULL: batch layout enabling the shared-memory staged path; max_graph_nodes /
 *   max_graph_edges: host-known maxima over the batch (0 = unknown -> generic path).
 *   complete_k: K if the caller vouches that the batch is G complete directed K-robot graphs without self loops in the
 *   reference's src-major edge order (dgl/dataloader.py:88-95; see rgnn_check_complete), else 0.  2 <= K <= 8 selects
 *   a register-tiled kernel with the topology compiled in. */
int rgnn_aggregate(int dtype, const void* x, int64_t ldx, const void* z, int use_beta,
                   const float* nbr_scale, int post, const float* dst_scale,
                   const void* addend, int64_t ldadd,
                   void* out, int64_t ldout, void* copy_x_to,
                   const int32_t* ptr32, const int32_t* nbr32, const int32_t* eid32,
                   const int32_t* node_offsets32, int64_t num_graphs,
                   int64_t max_graph_nodes, int64_t max_graph_edges, int64_t complete_k,
                   int64_t num_nodes, int64_t num_edges, int64_t feat, int64_t pix, void* stream);

/* ---- A10: per-edge FiLM gradients with the sigmoid derivative fused ----------------------
 * dz[e,2c]   = z[e,2c]  (1-z[e,2c])   * w_e * sum_p x[src_e,c,p] * dagg[dst_e,c,p]
 * dz[e,2c+1] = z[e,2c+1](1-z[e,2c+1]) * w_e * sum_p dagg[dst_e,c,p],     w_e = inv_in_deg[dst_e]
 * (autograd of models.py:211 and :150).  fuse_sigmoid_grad = 0 drops the z(1-z) factors (true d/dz).
 * src32/dst32 [E] in edge-id order. */
int rgnn_film_edge_grad(int dtype, const void* x, int64_t ldx, const void* dagg, int64_t lddagg,
                        const void* z, int fuse_sigmoid_grad, const float* inv_in_deg,
                        const int32_t* src32, const int32_t* dst32,
                        void* dz, int64_t num_edges, int64_t feat, int64_t pix, void* stream);

/* ---- A10 fused: dh AND dz of the FiLM-mean aggregation from one shared-memory staging of h and dAgg -----------
 * dh[u] = addend[u] + sum_{e: src_e = u} gamma_e * w_e * dagg[dst_e];  dz as in rgnn_film_edge_grad.
 * ptr32 / idx32 / node_offsets32 are the packed arrays rgnn_build_batched_csr wrote.  addend may be NULL or == dh.
 * Batched-robot-graph path only: returns RGNN_E_UNSUPPORTED when a graph does not fit shared memory or the rows
 * are not 16-byte addressable (callers then use rgnn_aggregate on the out-CSR + rgnn_film_edge_grad). */
int rgnn_film_backward(int dtype, const void* h, int64_t ldh, const void* dagg, int64_t lddagg, const void* z,
                       int fuse_sigmoid_grad, const void* addend, int64_t ldadd, void* dh, int64_t lddh, void* dz,
                       const float* inv_in_deg, const int32_t* ptr32, const int32_t* idx32, const int32_t* node_offsets32,
                       int64_t num_graphs, int64_t max_graph_nodes, int64_t max_graph_edges, int64_t complete_k,
                       int64_t num_nodes, int64_t num_edges, int64_t feat, int64_t pix, void* stream);

/* ---- A5/A6/A10 on RAGGED batches of complete graphs (samples with different camera counts, dgl/dataloader.py:88-99) ------
 * The host buckets the batch by robot count; each call covers ONE bucket of `bucket_graphs` complete `robots`-node graphs (2..8,
 * reference edge order) scattered through the batch: bases = int32 [bucket_graphs, 2] = (first node row, first edge row) of each.
 * Same arithmetic, layouts and results as rgnn_aggregate (FiLM, mean) / rgnn_film_backward on those graphs, with the topology
 * compiled in (streaming ring kernels) instead of a CSR walk.  RGNN_E_UNSUPPORTED when pix != 64 or the rows are not 16-byte
 * addressable (caller: the general kernels). */
int rgnn_film_bucket_forward(int dtype, const void* x, int64_t ldx, const void* z, void* out, int64_t ldout, void* copy_x_to,
                             const int32_t* bases, int64_t bucket_graphs, int64_t robots, int64_t feat, int64_t pix, void* stream);
int rgnn_film_bucket_backward(int dtype, const void* h, int64_t ldh, const void* dagg, int64_t lddagg, const void* z, int fuse_sigmoid_grad,
                              const void* addend, int64_t ldadd, void* dh, int64_t lddh, void* dz, const int32_t* bases,
                              int64_t bucket_graphs, int64_t robots, int64_t feat, int64_t pix, void* stream);

/* ---- A4 (+A10): first edge-encoder layer, K = 9 -----------------------------------------
 * a1 = relu(pose W1^T + b1)   (nn.Linear(9,F)+ReLU, dgl/model/models.py:147-148); pose [E,9] fp32
 * (edge.float(), models.py:153); w1 [F,9], b1 [F] fp32 master weights; a1 [E,F] in `dtype`. */
int rgnn_edge_mlp1_fwd(int dtype, const float* pose, const float* w1, const float* b1,
                       void* a1, int64_t num_edges, int64_t feat, void* stream);
/* dw1 [F,9], db1 [F] fp32 (overwritten) from da1 [E,F] (`dtype`, ReLU mask already applied). */
int rgnn_edge_mlp1_bwd(int dtype, const float* pose, const void* da1,
                       float* dw1, float* db1, int64_t num_edges, int64_t feat, void* stream);

/* ---- A4/A8 (+A10, A11): strided GEMM family on CUDA cores (fp32-exact products) -----------
 * For z in [0,batch):  C_z[m,n] = epi( sum_{r<red} sum_{k<K} A_z,r[m,k] * B_z,r[k,n] )
 * with element strides (sAz,sAr,sAm,sAk), (sBz,sBr,sBk,sBn), (sCz,sCm,sCn); aux shares C's strides.
 * split > 1: the `red` loop (or, when red == 1, the K loop) is cut into `split` chunks, partial sums go to
 * `partial` ([split, M, N] fp32, batch must be 1) and are combined in a fixed order (deterministic), then `epi`.
 * a_dtype/b_dtype/c_dtype: RGNN_F32 | RGNN_BF16; bias fp32.
 * Covers nn.Linear(F,2F)+Sigmoid (models.py:149-150), Conv2d(2F,F,1) (models.py:164,184) and the
 * dX / dW products of their autograd, plus the GraphConv / GCN2Conv weight transforms. */
int rgnn_gemm(int a_dtype, const void* A, int64_t sAz, int64_t sAr, int64_t sAm, int64_t sAk,
              int b_dtype, const void* B, int64_t sBz, int64_t sBr, int64_t sBk, int64_t sBn,
              int c_dtype, void* C, int64_t sCz, int64_t sCm, int64_t sCn,
              int64_t M, int64_t N, int64_t K, int64_t batch, int64_t red, int64_t split, float* partial,
              int epi, const float* bias, const void* aux, float alpha, float beta, void* stream);

/* ---- A4/A8 (+A10): the same transforms on the tcgen05 tensor cores (bf16 in, fp32 accumulate in TMEM) -------
 * Throughput path.  Shapes the tiles do not cover return RGNN_E_UNSUPPORTED and the caller uses rgnn_gemm. */

/* fp32 master weights [rows, cols] -> bf16 copy and/or bf16 transpose [cols, rows] (either may be NULL) */
int rgnn_tc_cast_weights(const float* w, int64_t rows, int64_t cols, void* out_bf16, void* out_t_bf16, void* stream);

/* Same cast of alpha * w + ident * I (square w when ident != 0): GCN2Conv's identity mapping (1-beta) I + beta W folded
 * into the bf16 weight operand, so the transform is ONE plain GEMM r . W' with no second operand in its epilogue. */
int rgnn_tc_cast_weights_affine(const float* w, int64_t rows, int64_t cols, float alpha, float ident, void* out_bf16, void* out_t_bf16,
                                void* stream);

/* 1x1 convolution over NCHW robot feature maps with P = 64 pixels, concat-free (models.py:182-184 and its dX):
 * out[v,m,:] = sum_{k<c0} W[m,k] x0[v,k,:] + sum_{k<c1} W[m,c0+k] x1[v,k,:] (+ bias[m]);  W bf16 [m0+m1, c0+c1];
 * rows m < m0 land in out0 [nodes, m0, 64], rows >= m0 in out1 [nodes, m1, 64] (x1/out1 optional with c1/m1 = 0). */
int rgnn_tc_conv1x1(const void* W, const void* x0, int64_t c0, const void* x1, int64_t c1,
                    void* out0, int64_t m0, void* out1, int64_t m1, const float* bias, int64_t nodes, void* stream);

/* the two steps above FUSED (north_star (c)): FiLM-mean aggregation in the prologue of the compress GEMM.
 * y[v] = W[:, :F] h[v] + W[:, F:] g[v] + bias with g = mean_{u != v} (gamma_{u->v} h[u] + beta_{u->v}) over `graphs` complete
 * `robots`-node graphs (2..8, reference edge order); h [G*K, F, 64], z [G*K(K-1), 2F] interleaved (gamma,beta), W [Fo, 2F],
 * y [G*K, Fo, 64], all bf16; bias fp32 [Fo] or NULL.  g is built in shared memory in the tensor-core operand layout and is
 * written to g_out only if g_out != NULL (the backward needs it); otherwise it never reaches HBM.  F %% 32 == 0, Fo %% 128 == 0. */
int rgnn_tc_conv_film_fused(const void* W, const void* h, const void* z, const float* bias, void* y, void* g_out,
                            int64_t graphs, int64_t robots, int64_t F, int64_t Fo, void* stream);

/* north_star (c)/(d), backward: the FiLM-mean backward FUSED INTO THE EPILOGUE of the compress conv's dX GEMM - the gradient of the
 * concat, dcat = [dh_direct ; dg_h], stays in TMEM and never reaches HBM (autograd of models.py:181-184, 207-211):
 *   dcat_v = Wc^T dY_v;  dh[u] = dcat[u,:F] + 1/(K-1) sum_{a != u} gamma_{u->a} dcat[a,F:];  dgamma_{u->a} = 1/(K-1) sum_p h[u,:,p] dcat[a,F:,p];
 *   dbeta_{u->a} = 1/(K-1) sum_p dcat[a,F:,p];  dz = (dgamma, dbeta) [* sigmoid' when fuse_sigmoid_grad].
 * W_perm: bf16 [2F, Fo] = Wc^T with rows permuted per 64-channel tile t, j < 8: rows 128t+16j+(0..7) = direct channels 64t+8j+(0..7), rows
 * 128t+16j+8+(0..7) = the aggregated rows F+64t+8j+(0..7) (a channel's two accumulator rows share a warp).  dy [G K, Fo, 64], h [G K, F, 64],
 * z / dz [G K(K-1), 2F] interleaved (gamma,beta), dh [G K, F, 64], all bf16; G complete `robots`-node graphs (2..6), reference edge order.
 * HBM traffic: dY + h in, dh out (3 tensor passes) instead of 7 for rgnn_tc_conv1x1 (dX) + rgnn_film_backward. */
int rgnn_tc_conv_dx_film_bwd(const void* W_perm, const void* dy, const void* h, const void* z, int fuse_sigmoid_grad,
                             void* dh, void* dz, float* dbias2, int64_t graphs, int64_t robots, int64_t F, int64_t Fo, void* stream);
/* dbias2 (nullable): fp32 [2F]; receives the column sums of dz (= the bias gradient of nn.Linear(F,2F), models.py:149); overwritten. */

/* weight gradient of the above: dW[m,k] = sum_v sum_p dy[v,m,p] x[v,k,p], x = x0 | x1; dW fp32 (row stride ldw);
 * dbias (nullable) fp32 [M] = sum_v sum_p dy[v,m,p], computed by the otherwise idle epilogue warps from the dy tiles the
 * mainloop stages anyway (no second pass over dy); partial fp32 [splits, M, c0+c1 (+1 if dbias)] scratch; the robots are
 * split across CTAs and summed in a fixed order (deterministic). */
int rgnn_tc_conv1x1_wgrad(const void* dy, int64_t M, const void* x0, int64_t c0, const void* x1, int64_t c1,
                          float* dW, int64_t ldw, float* dbias, float* partial, int64_t splits, int64_t nodes, void* stream);

/* C[M,N] = epi(alpha * A[M,K] . B[N,K]^T), bf16 row-major operands; epi = bit-or of RGNN_EPI_BIAS_N | _SIGMOID | _RELU_MASK |
 * _ADD_AUX (+ beta * aux[m,n], aux bf16 with C's row stride).  nn.Linear(F,2F)+Sigmoid (models.py:149-150) and dA1 of its
 * autograd; the GraphConv weight transform; GCN2Conv's identity mapping (1-b) X + b X W fused in the epilogue. */
int rgnn_tc_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                    int64_t M, int64_t N, int64_t K, int epi, const float* bias, const void* aux, float alpha, float beta,
                    void* stream);

/* C[M,N] fp32 = alpha * A^T . B with A bf16 [K,M], B bf16 [K,N] (dW2 = dZ^T . A1; dW of GraphConv/GCN2Conv); split over K,
 * partial fp32 [splits,M,N]. */
int rgnn_tc_gemm_nn(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                    int64_t M, int64_t N, int64_t K, float alpha, float* partial, int64_t splits, void* stream);

/* ---- the fp32 PARITY mode on the tensor cores: split-bf16 products (SURVEY §7 "fp32 parity vs tensor cores") ----------------
 * An fp32 operand x is stored as three bf16 planes with x = x0 + x1 + x2 (8 significand bits each).  A.B is then the sum of the six
 * plane-pair products A_i.B_j, i + j <= 2 (the three dropped ones are ~2^-24 relative), run as ONE bf16 tcgen05 GEMM whose k-loop
 * walks the k-blocks six times - 2^-16 terms first, A_0.B_0 last - into one fp32 TMEM accumulator; outputs are fp32.  Measured
 * against fp64 (profiles/r02_fp32_split_study.jsonl): <= 2.5e-6 normalised max error up to K = 2560 per accumulator (cuBLAS fp32 on
 * CUDA cores: 1.7e-6; one bf16 pass: 2.5e-3; one TF32 pass: 3e-4), so the reference's fp32 nn.Linear(F,2F)+Sigmoid (models.py:149-150),
 * Conv2d(2F,F,1) (models.py:164) and their autograd stay inside the 1e-5 contract.  Reductions longer than ~4096 (the weight
 * gradients) must be split (`splits`): the accumulator truncates when a small addend is aligned to a large running sum. */

/* x fp32 [n] -> out bf16 [3, plane]: planes x0, x1, x2; elements [n, plane) of every plane are zeroed.  n %% 4 == 0, plane %% 8 == 0. */
int rgnn_tc_split3(const float* x, int64_t n, int64_t plane, void* out_bf16, void* stream);

/* fp32 master weights [rows, cols] -> planes [3, rows, cols] and/or planes of the transpose [3, cols, rows] (either may be NULL) */
int rgnn_tc_split3_weights(const float* w, int64_t rows, int64_t cols, void* out3, void* out3_t, void* stream);

/* rgnn_tc_conv1x1 on split planes: W3 [3, m0+m1, c0+c1]; x0_3 [3, plane_nodes, c0, 64], x1_3 [3, plane_nodes, c1, 64] (plane_nodes >= nodes);
 * out0 / out1 fp32 [nodes, m, 64].  (models.py:182-184 in fp32 and its dX.) */
int rgnn_tc3_conv1x1(const void* W3, const void* x0_3, int64_t c0, const void* x1_3, int64_t c1, int64_t plane_nodes,
                     float* out0, int64_t m0, float* out1, int64_t m1, const float* bias, int64_t nodes, void* stream);

/* rgnn_tc_conv1x1_wgrad on split planes: dy3 [3, plane_nodes, M, 64], x0_3 / x1_3 as above; dW fp32.  Choose splits >= nodes / 64 so that
 * one accumulator sums at most 64 robots x 64 pixels.  (The bias gradient is a separate rgnn_reduce_sum in this mode.) */
int rgnn_tc3_conv1x1_wgrad(const void* dy3, int64_t M, const void* x0_3, int64_t c0, const void* x1_3, int64_t c1, int64_t plane_nodes,
                           float* dW, int64_t ldw, float* partial, int64_t splits, int64_t nodes, void* stream);

/* rgnn_tc_gemm_tn on split planes: A3 [3, plane_a, lda] (plane_a >= M rows), B3 [3, plane_b, ldb] (plane_b >= N); C fp32 [M, ldc];
 * epi = bit-or of RGNN_EPI_BIAS_N | RGNN_EPI_SIGMOID (exact exp-based sigmoid).  C = epi(alpha * A . B^T). */
int rgnn_tc3_gemm_tn(const void* A3, int64_t lda, int64_t plane_a, const void* B3, int64_t ldb, int64_t plane_b, float* C, int64_t ldc,
                     int64_t M, int64_t N, int64_t K, int epi, const float* bias, float alpha, void* stream);

/* rgnn_tc_gemm_nn on split planes: A3 [3, plane_a, lda], B3 [3, plane_b, ldb] with plane rows = K padded to a multiple of 64 (zero rows). */
int rgnn_tc3_gemm_nn(const void* A3, int64_t lda, int64_t plane_a, const void* B3, int64_t ldb, int64_t plane_b, float* C, int64_t ldc,
                     int64_t M, int64_t N, int64_t K, float alpha, float* partial, int64_t splits, void* stream);

/* out = a > 0 ? da : 0 elementwise over n values (ReLU backward, models.py:148; the bf16 path fuses it into the dA1 GEMM epilogue) */
int rgnn_relu_grad(int dtype, const void* a, const void* da, void* out, int64_t n, void* stream);

/* ---- encoder -> GNN and GNN -> decoder seams (models.py:177-180 hand-over, :190-205 / blocks.py:13-25 decoder) ----
 * out[v, c, r] = (out_dtype) x[v, r, c]: per-robot transpose [rows, cols] -> [cols, rows] with the dtype conversion fused, one pass.
 * channels-last backbone output [N, H*W, F] -> the NCHW [N, F, H*W] the robot-graph kernels consume: rows = H*W, cols = F;
 * NCHW block output -> channels-last for the decoder's first conv: rows = F, cols = H*W. */
int rgnn_layout_cast(int in_dtype, const void* x, int out_dtype, void* out, int64_t nodes, int64_t rows, int64_t cols, void* stream);

/* dzpre = dz * z * (1 - z)  elementwise over n values (sigmoid backward, models.py:150) */
int rgnn_sigmoid_grad(int dtype, const void* z, const void* dz, void* dzpre, int64_t n, void* stream);

/* out = a*x + b*y elementwise over n values (GCN2Conv initial residual (1-alpha)*agg + alpha*feat_0) */
int rgnn_axpby(int dtype, float a, const void* x, float b, const void* y, void* out, int64_t n, void* stream);

/* GCN2Conv on complete K-robot graphs (D^-1/2 A D^-1/2 of a complete graph is the mean over the other K-1
 * robots): out[v] = sa * mean_{u != v, same graph} x[u] + sb * addend[v] in one streaming pass (addend may be NULL).
 * Rows of `row` elements per node (node strides ldx / ldadd / ldout); node b*K+i is robot i of graph b (dataloader.py:88-95).
 * Returns RGNN_E_UNSUPPORTED when robots > 8 or the row is not a multiple of 256 16-byte vectors (caller: rgnn_aggregate + rgnn_axpby). */
int rgnn_complete_mean_mix(int dtype, const void* x, int64_t ldx, const void* addend, int64_t ldadd, void* out, int64_t ldout,
                           int64_t num_graphs, int64_t robots, int64_t row, float sa, float sb, void* stream);

/* out[r,c] = x[r,c] + bias[c]  (GraphConv bias after a plain aggregation) */
int rgnn_bias_add(int dtype, const void* x, const float* bias, void* out, int64_t rows, int64_t cols, void* stream);

/* out[j] = sum_{i<outer} sum_{k<inner} x[i*s_outer + j*s_j + k*s_inner]   (bias gradients: fp32 out) */
int rgnn_reduce_sum(int dtype, const void* x, int64_t outer, int64_t s_outer, int64_t len, int64_t s_j,
                    int64_t inner, int64_t s_inner, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ROBOT_GNN_H_ */
