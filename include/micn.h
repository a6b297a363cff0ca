/*
 * micn.h - C ABI of the B200-native modality-conditioned instance norm ("instance_cond") for MI-Seg.
 *
 * This is the drop-in boundary for the one hot path this repo accelerates.  Each entry point
 * replaces a piece of the reference's Python/ATen path (paths relative to the MI-Seg checkout):
 *
 *   micn_fwd   <- networks/norms/conditional_instance_norm.py:59-60  (_apply_instance_norm: per-sample
 *                 nn.InstanceNorm{1,2,3}d picked by styles[i] + torch.stack), :52-57 (un-batched input),
 *                 and, with epilogue != MICN_EPI_NONE, the activation / residual that follows it in
 *                 networks/blocks/dynunet_block.py:107-111, :113-125 (UnetResBlock) and :188-202
 *                 (UnetBasicBlock): LeakyReLU(0.01) and "out += residual".
 *   micn_bwd   <- the autograd graph of the above (StackBackward -> native_batch_norm_backward ->
 *                 RepeatBackward -> AccumulateGrad on norms[s].weight / norms[s].bias).
 *
 * Conventions
 *   - plain C types only; every pointer marked "device" is a CUDA device pointer valid on the
 *     device that is current when the call is made; `stream` is a cudaStream_t passed as void*.
 *   - the calls only enqueue work on `stream`: no allocation, no host synchronisation, no
 *     exceptions.  Return 0 on success, a MICN_ERR_* (< 0) for argument errors, or a positive
 *     cudaError_t from the launch.  micn_error_string() decodes either.
 *   - a "slab" is the D*H*W (= M) voxels of one (sample n, channel c); x is addressed as
 *     x[n*x_stride_n + c*x_stride_c + m] (strides in ELEMENTS, slab itself dense);
 *     y, residual, act_out, dy, dx, dresidual are dense [N, C, M].
 *   - statistics, gamma/beta and all accumulations are fp32 regardless of the I/O dtype.
 *   - gamma/beta are HOST arrays of `num_styles` DEVICE pointers (norms[s].weight / .bias of the
 *     reference's ModuleList); pass NULL for a non-affine norm (gamma = 1, beta = 0).
 *   - styles is a DEVICE int64 array [N] read by the kernels (no host sync); python-style negative
 *     indices wrap; an out-of-range style is clamped and recorded in the workspace status word
 *     (micn_read_status) instead of faulting.  styles == NULL means style 0 for every sample.
 */
#ifndef MICN_H_
#define MICN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MICN_VERSION 100 /* 0.1.0 */
#define MICN_MAX_STYLES 16
#define MICN_MAX_PEERS 16 /* GPUs of one NVLink box taking part in micn_bwd_allreduce */

/* I/O element types */
enum { MICN_F32 = 0, MICN_BF16 = 1, MICN_F16 = 2 };

/* fused epilogues (dynunet_block.py) */
enum {
    MICN_EPI_NONE = 0,      /* y = norm(x)                                  */
    MICN_EPI_LRELU = 1,     /* y = lrelu(norm(x))            :107-111       */
    MICN_EPI_ADD_LRELU = 2, /* y = lrelu(norm(x) + residual) :113-125       */
    MICN_EPI_NORM_ADD_LRELU = 3 /* y = lrelu(norm_a(a) + norm_b(b)): the downsample branch, :113-125 with conv3 /
                                   norm3 (:82-98); micn_fwd_dual / micn_bwd_dual only */
};

/* error codes (negative; positive values are cudaError_t) */
enum {
    MICN_OK = 0,
    MICN_ERR_BAD_ARG = -1,
    MICN_ERR_BAD_DTYPE = -2,
    MICN_ERR_TOO_MANY_STYLES = -3,
    MICN_ERR_WORKSPACE = -4,
    MICN_ERR_UNALIGNED = -5,
    MICN_ERR_NO_DEVICE = -6,
    MICN_ERR_UNSUPPORTED = -7 /* micn_fwd_dual / micn_bwd_dual: no dual kernel takes this shape */
};

int micn_version(void);
const char* micn_error_string(int code);

/* Tuning / experiment knobs ("force_path" 0 small / 1 cluster / 2 flat / 4 resident, "flat_slots", "flat_lag",
 * "flat_piece_vecs", "cluster_size", ...) and read-backs ("last_path", "launches").  Returns 0 or
 * MICN_ERR_BAD_ARG for an unknown key.  value < 0 restores the automatic choice.
 *
 * Launch modes.  Every kernel starts with griddepcontrol.wait; the small, resident and fused channels-last kernels are
 * launched as programmatic dependents (their launch latency overlaps the tail of the kernel before them in the stream;
 * "pdl" = 0 switches that off).  The persistent flat kernels - whose CTAs wait for each other's records across the grid -
 * are launched COOPERATIVELY by default, which also keeps two such kernels launched from different streams from
 * dead-locking each other on half a GPU each.  A process that issues all its kernels on ONE stream (the usual training
 * loop) can set "flat_pdl" = 1 and "flat_coop" = 0: the flat kernels then queue as programmatic dependents too, which
 * saves about 2 us per launch (1 x 48 x 96^3 bf16 forward + backward: 80.1 -> 76.6 us). */
int micn_set_option(const char* key, long long value);
long long micn_get_option(const char* key);

/* Bytes of device workspace micn_fwd/micn_bwd need for this problem (cross-CTA exchange records of
 * the flat path, per-slab sums).  The workspace must be zero-filled ONCE when it is allocated and can
 * then be reused by any number of calls of any shape that fits: exchange records are tagged with a
 * launch epoch that lives in the workspace header and is advanced by the kernels themselves, so the calls
 * are safe to capture in a CUDA graph and replay (no per-launch state on the host).  Calls that may
 * overlap in time (different streams) need different workspaces.  The pointer must be 16-byte aligned
 * (MICN_ERR_UNALIGNED otherwise).  Without a workspace (or with one that is too small) the large-slab fast
 * path is not used. */
size_t micn_workspace_bytes(int64_t N, int64_t C, int64_t M, int dtype, int num_styles);

/* Host-blocking read (cudaMemcpy) of the sticky status word: bit 0 = a style index was out of
 * range.  Clears the word.  Debug aid; never called on the hot path. */
int micn_read_status(void* workspace, void* stream, int* status_out);

/* Forward.  save_mean / save_rstd: device fp32 [N*C] (needed by micn_bwd; may be NULL for
 * inference).  residual: device, dense [N,C,M], only for MICN_EPI_ADD_LRELU. */
int micn_fwd(const void* x, void* y, const void* residual,
             const float* const* gamma, const float* const* beta, int num_styles,
             const int64_t* styles,
             float* save_mean, float* save_rstd,
             int64_t N, int64_t C, int64_t M,
             int64_t x_stride_n, int64_t x_stride_c,
             int dtype, int epilogue, float slope, float eps,
             void* workspace, size_t workspace_bytes, void* stream);

/* Backward.  dy is the gradient of the (post-epilogue) output.  act_out is the forward OUTPUT
 * tensor (needed only for MICN_EPI_ADD_LRELU, to recover the LeakyReLU mask); for MICN_EPI_LRELU the
 * mask is recomputed from x and the saved statistics.  dresidual (ADD_LRELU only) receives the
 * gradient of `residual`.  dgamma / dbeta: device fp32 [num_styles*C], OVERWRITTEN with the sums over
 * the samples of each style (zero rows for styles absent from the batch); either both or neither
 * may be NULL. */
int micn_bwd(const void* dy, const void* x, const void* act_out,
             const float* const* gamma, const float* const* beta, int num_styles,
             const int64_t* styles,
             const float* save_mean, const float* save_rstd,
             void* dx, void* dresidual,
             float* dgamma, float* dbeta,
             int64_t N, int64_t C, int64_t M,
             int64_t x_stride_n, int64_t x_stride_c,
             int dtype, int epilogue, float slope,
             void* workspace, size_t workspace_bytes, void* stream);

/* Backward with the data-parallel exchange of the per-style parameter gradients FUSED into the kernel (the
 * one collective of the path; tune.py:103-109 does it with DDP's NCCL all-reduce).  One process per GPU of an NVLink box;
 * `peer_bufs` is a HOST array of `world` DEVICE pointers: entry r is rank r's exchange buffer as mapped into THIS process
 * (CUDA IPC / symmetric memory; entry `rank` is the local one), each micn_peer_buffer_bytes() large, 16-byte aligned and
 * zero-filled once.  The kernel keeps the final sums of every channel as self-validating 16-byte records in its own buffer
 * and, once its gather warps are through their pieces, stores them straight into every peer's buffer (st.relaxed.sys over
 * NVLink, hidden behind the last normalise tasks); the `world` records of every channel are folded from the rank's OWN
 * buffer in rank order into dgamma / dbeta = the SUM over all ranks (bit-identical on every rank):
 *   MICN_FOLD_THIS      at the end of this kernel (synchronous all-reduce: exposes one NVLink latency + the skew between
 *                       the GPUs, a few microseconds, instead of a ~17 us NCCL launch);
 *   MICN_FOLD_PREVIOUS  at the end of this kernel too, but the PREVIOUS call's exchange (its records arrived a whole kernel ago:
 *                       nothing to wait for) - dgamma / dbeta then lag one call behind, the way an asynchronous bucketed
 *                       all-reduce completes behind the backward pass; micn_allreduce_fold() delivers the last call's;
 *   MICN_FOLD_NONE      only store the records (fold later with micn_allreduce_fold()).
 * Requirements: every rank makes the same sequence of micn_bwd_allreduce calls on its buffer (the record tag is a launch
 * counter kept in the buffer, CUDA-graph safe); the problem must take the flat path (MICN_ERR_UNSUPPORTED otherwise: fall
 * back to micn_bwd + NCCL). */
enum { MICN_FOLD_NONE = 0, MICN_FOLD_THIS = 1, MICN_FOLD_PREVIOUS = 2 };
size_t micn_peer_buffer_bytes(int64_t C, int num_styles, int world);
/* Folds the exchange of the most recent micn_bwd_allreduce call on this buffer (a small kernel on `stream`). */
int micn_allreduce_fold(void* const* peer_bufs, int rank, int world, int64_t C, int num_styles, float* dgamma, float* dbeta,
                        void* stream);
int micn_bwd_allreduce(const void* dy, const void* x, const void* act_out,
                       const float* const* gamma, const float* const* beta, int num_styles,
                       const int64_t* styles,
                       const float* save_mean, const float* save_rstd,
                       void* dx, void* dresidual,
                       float* dgamma, float* dbeta,
                       int64_t N, int64_t C, int64_t M,
                       int64_t x_stride_n, int64_t x_stride_c,
                       int dtype, int epilogue, float slope,
                       void* workspace, size_t workspace_bytes,
                       void* const* peer_bufs, int rank, int world, int fold_mode, void* stream);

/* The same two calls with the activation slope read from DEVICE memory: `slope_dev` points at the single weight of
 * the nn.PReLU that follows the norm in C-UNet's ADN block ("NDA": norm -> dropout(0) -> PReLU,
 * networks/blocks/acti_norm.py:104-110, convolutions.py:173-179), so prelu(norm(x)) and its input gradient are one
 * kernel each and no host read of the parameter is needed.  epilogue must be MICN_EPI_LRELU or MICN_EPI_ADD_LRELU; with
 * MICN_EPI_ADD_LRELU the backward recovers the activation mask from the sign of the OUTPUT, which is only valid for a
 * slope > 0, and no slope gradient is produced (the Python module refuses that combination: MI-Seg has no
 * prelu(norm(x) + r) - C-UNet adds its residual after the activation, convolutions.py:329).
 * The gradient of the slope itself is sum over pre <= 0 of dy * pre: for MICN_EPI_LRELU micn_bwd_prelu accumulates it in
 * the same pass and writes PARTIAL sums into `dslope_partial` (device fp32, at least max(N*C, 1024) entries,
 * ZERO-FILLED by the caller; one entry per CTA or per slab is written, the total is the gradient; may be NULL). */
int micn_fwd_prelu(const void* x, void* y, const void* residual,
                   const float* const* gamma, const float* const* beta, int num_styles,
                   const int64_t* styles,
                   float* save_mean, float* save_rstd,
                   int64_t N, int64_t C, int64_t M,
                   int64_t x_stride_n, int64_t x_stride_c,
                   int dtype, int epilogue, const float* slope_dev, float eps,
                   void* workspace, size_t workspace_bytes, void* stream);
int micn_bwd_prelu(const void* dy, const void* x, const void* act_out,
                   const float* const* gamma, const float* const* beta, int num_styles,
                   const int64_t* styles,
                   const float* save_mean, const float* save_rstd,
                   void* dx, void* dresidual,
                   float* dgamma, float* dbeta,
                   int64_t N, int64_t C, int64_t M,
                   int64_t x_stride_n, int64_t x_stride_c,
                   int dtype, int epilogue, const float* slope_dev,
                   float* dslope_partial,
                   void* workspace, size_t workspace_bytes, void* stream);

/* The downsample branch of UnetResBlock in one pass: y = lrelu(norm_a(a) + norm_b(b)) with a = conv2's output under
 * norm2 and b = conv3's output under norm3 (networks/blocks/dynunet_block.py:113-125 with :82-98) - C-Swin-UNETR's and
 * C-UNETR's `encoder1` and every decoder block take it.  Both tensors are dense [N, C, M]; both norms share `styles`
 * and `num_styles` (pass 1 and NULL styles for plain instance norms).  Forward reads a and b once and writes y once
 * (3*E*s bytes instead of 2 + 3 for norm3 followed by the add_lrelu call); backward reads a, b, dy once and writes da, db
 * (5*E*s instead of 4 + 3), recomputing the LeakyReLU mask from the saved statistics, and fills the parameter
 * gradients of both norms (dbeta_b equals dbeta_a by construction; either pair may be NULL).
 * micn_dual_supported() says whether these entry points take a problem of that shape (they return
 * MICN_ERR_UNSUPPORTED otherwise and the caller composes micn_fwd(b) + micn_fwd(a, ADD_LRELU) instead). */
int micn_dual_supported(int64_t N, int64_t C, int64_t M, int dtype, int backward);
int micn_fwd_dual(const void* a, const void* b, void* y,
                  const float* const* gamma_a, const float* const* beta_a,
                  const float* const* gamma_b, const float* const* beta_b, int num_styles,
                  const int64_t* styles,
                  float* save_mean_a, float* save_rstd_a, float* save_mean_b, float* save_rstd_b,
                  int64_t N, int64_t C, int64_t M, int dtype, float slope, float eps,
                  void* workspace, size_t workspace_bytes, void* stream);
int micn_bwd_dual(const void* dy, const void* a, const void* b,
                  const float* const* gamma_a, const float* const* beta_a,
                  const float* const* gamma_b, const float* const* beta_b, int num_styles,
                  const int64_t* styles,
                  const float* save_mean_a, const float* save_rstd_a, const float* save_mean_b, const float* save_rstd_b,
                  void* da, void* db,
                  float* dgamma_a, float* dbeta_a, float* dgamma_b, float* dbeta_b,
                  int64_t N, int64_t C, int64_t M, int dtype, float slope,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Channels-last (token-major) variant: x, y, dy, dx are dense [N, M, C] tensors (C fastest, C even) - the layout in
 * which PatchMerging's norms (networks/blocks/patch_merging.py:136-141) and the ViT token norms
 * (transformer_block.py:87-92, vit.py:188-193) reach `_apply_instance_norm` as permuted views.  Same math, no
 * epilogue; the output keeps the input's layout, so the transposing copies around the norm disappear.
 * Workspace: micn_cl_workspace_bytes, zero-filled once when allocated (16-byte aligned), then reusable forever. */
size_t micn_cl_workspace_bytes(int64_t N, int64_t C, int64_t M);
int micn_fwd_cl(const void* x, void* y,
                const float* const* gamma, const float* const* beta, int num_styles,
                const int64_t* styles,
                float* save_mean, float* save_rstd,
                int64_t N, int64_t C, int64_t M, int dtype, float eps,
                void* workspace, size_t workspace_bytes, void* stream);
int micn_bwd_cl(const void* dy, const void* x,
                const float* const* gamma, const float* const* beta, int num_styles,
                const int64_t* styles,
                const float* save_mean, const float* save_rstd,
                void* dx, float* dgamma, float* dbeta,
                int64_t N, int64_t C, int64_t M, int dtype,
                void* workspace, size_t workspace_bytes, void* stream);

/* Host-buffer convenience path: x (and dy) live in HOST memory (pinned for full speed); the call
 * stages slab groups through `dev_scratch` (device, >= micn_host_scratch_bytes) with
 * H2D / kernels / D2H overlapped on internal streams, and BLOCKS until y (and dx, dgamma, dbeta)
 * are back in host memory.  This is what `e2e` in bench.py times. */
size_t micn_host_scratch_bytes(int64_t N, int64_t C, int64_t M, int dtype, int num_styles, int with_backward);
int micn_fwd_bwd_host(const void* x_host, const void* dy_host, void* y_host, void* dx_host,
                      const float* gamma_host, const float* beta_host, int num_styles, /* [S*C] host */
                      const int64_t* styles_host,
                      float* dgamma_host, float* dbeta_host,                           /* [S*C] host */
                      int64_t N, int64_t C, int64_t M, int dtype, int epilogue, float slope, float eps,
                      void* dev_scratch, size_t dev_scratch_bytes);

#ifdef __cplusplus
}
#endif
#endif /* MICN_H_ */
