/* icp_b200.h -- C ABI of libicp_b200.so: the B200-native (sm_100a) photogeometric ICP hot path.
 *
 * This is the drop-in boundary for the reference's per-iteration ICP path (nlamprian/ICP).  Every entry
 * point below replaces one host-class operation of /root/reference/include/ICP/algorithms.hpp (cited per
 * function as algorithms.hpp:LINE / algorithms.cpp:LINE = src/ICP/algorithms.cpp); the header-only C++
 * classes of the same names in include/ICP/algorithms.hpp call these functions and nothing else.
 *
 * Conventions
 *   - plain C types only; `void*` / `float*` arguments named d_* are DEVICE pointers (cudaMalloc'd, e.g. by
 *     icp_malloc or by torch); h_* are host pointers.  No torch / OpenCL types anywhere.
 *   - a point is 8 x f32 `[x y z 1 r g b 1]` (pc8d); a pose is 8 x f32 `{qx qy qz qw tx ty tz s}` (D_IO_T,
 *     algorithms.hpp:2245-2254).
 *   - every call enqueues on the context's stream and returns without synchronising unless stated.
 *   - return value: 0 = ok, non-zero = error; icp_last_error() gives the message (thread local).
 *     ICP_ERR_CONFIG mirrors the reference's "Error[<Class>]: ..." + exit (algorithms.cpp:164-168);
 *     ICP_ERR_CUDA mirrors a thrown cl::Error.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with ICP_ERR_CUDA.
 */
#ifndef ICP_B200_H
#define ICP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICP_OK          0
#define ICP_ERR_CONFIG  1
#define ICP_ERR_CUDA    2
#define ICP_ERR_ARG     3

typedef struct icp_ctx  icp_ctx;   /* device ordinal + stream + scratch arena (replaces clutils::CLEnv + CLEnvInfo<1>) */
typedef struct icp_step icp_step;  /* one registration engine (replaces ICPStep<CR,CW> / ICP<CR,CW>) */
typedef struct icp_batch icp_batch;/* many independent registrations per launch (throughput mode) */

/* rbc_dist_id of RBC/data_types.hpp, layout stated at kernels/icp_kernels.cl:34-38 */
typedef struct { float dist; uint32_t id; } icp_dist_id;

/* ICPStepConfigT / ICPStepConfigW  (algorithms.hpp:1544-1564) */
enum { ICP_ROT_EIGEN = 0, ICP_ROT_POWER_METHOD = 1 };
enum { ICP_W_REGULAR = 0, ICP_W_WEIGHTED = 1 };
/* engine execution modes: same results bit for bit */
enum { ICP_MODE_STAGED = 0,   /* one kernel per reference stage (what run(timer) times, algorithms.hpp:2359-2399) */
       ICP_MODE_FUSED  = 1 }; /* fused kernels, CUDA-graph replay, no host round trip */

const char *icp_last_error(void);
const char *icp_version(void);

/* ---- context (replaces clutils::CLEnv::addContext/addQueue; one in-order queue, algorithms.cpp:57) ---- */
int  icp_ctx_create(int device, void *cuda_stream /* cudaStream_t or NULL = own stream */, icp_ctx **out);
void icp_ctx_destroy(icp_ctx *ctx);
int  icp_ctx_sync(icp_ctx *ctx);                               /* queue.finish () */
int  icp_device_info(icp_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, int *clock_khz, size_t *l2_bytes);

/* ---- memory (replaces cl::Buffer creation, enqueueWriteBuffer / enqueueReadBuffer) ---- */
int  icp_malloc(icp_ctx *ctx, size_t bytes, void **d_ptr);
int  icp_free(icp_ctx *ctx, void *d_ptr);
int  icp_host_alloc(size_t bytes, void **h_ptr);               /* pinned; CL_MEM_ALLOC_HOST_PTR staging */
int  icp_host_free(void *h_ptr);
int  icp_memcpy_h2d(icp_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, int block);
int  icp_memcpy_d2h(icp_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, int block);
int  icp_memcpy_d2d(icp_ctx *ctx, void *d_dst, const void *d_src, size_t bytes);
int  icp_memset(icp_ctx *ctx, void *d_dst, int value, size_t bytes);

/* ---- timing on the context stream (replaces clutils::GPUTimer, algorithms.hpp:140-163) ---- */
int  icp_timer_start(icp_ctx *ctx);
int  icp_timer_stop(icp_ctx *ctx, float *ms);                  /* synchronises on the stop event */
int  icp_flush_l2(icp_ctx *ctx);                               /* overwrite a buffer larger than L2 (benchmark hygiene) */

/* ---- pipeline stages (one call = one reference stage class ::run()) ---- */
/* ICPLMs::run, algorithms.hpp:312-374, algorithms.cpp:621-786, kernel icp_kernels.cl:62-76 */
int icp_get_lms(icp_ctx *ctx, const float *d_cloud /*640*480*8*/, float *d_lms /*16384*8*/);
/* RGB-D frame -> pc8d cloud: the conversion the reference's frame grabber applies before writing kg_pc8d_*.bin
 * (src/kinect_frame_grabber.cpp:246-263; RGBDTo8D::init (640, 480, 595.f, 1.f, ...) at :222 for the filtered path).
 * depth: W*H uint16 (mm, 0 = invalid), rgb: W*H*3 uint8, out: W*H*8 f32 [x y z 1 r g b 1]. */
int icp_rgbd_to_pc8d(icp_ctx *ctx, const uint16_t *d_depth, const uint8_t *d_rgb, uint32_t W, uint32_t H, float focal, float *d_cloud);
/* ICPReps::run, algorithms.hpp:397-459, algorithms.cpp:791-977, kernel :96-114; generalised to a W x H grid */
int icp_get_reps(icp_ctx *ctx, const float *d_lms, uint32_t W, uint32_t H, uint32_t nr, float *d_reps);
/* ICPTransform<QUATERNION>::run, algorithms.hpp:1239-1320, kernel :771-802 */
int icp_transform_quaternion(icp_ctx *ctx, const float *d_M, const float *d_T8, float *d_out, uint32_t m);
/* ICPTransform<MATRIX>::run, algorithms.hpp:1347-1424, kernel :903-933 (T row-major 4x4) */
int icp_transform_matrix(icp_ctx *ctx, const float *d_M, const float *d_T16, float *d_out, uint32_t m);
/* RBC::RBCConstruct<KINECT_R,GENERIC>::run as wired at algorithms.cpp:4503-4508, :4658-4659.
 * Outputs: d_rep_id[n] (nearest representative of every point), d_N[nr], d_O[nr] (list sizes / offsets),
 * d_perm[n] (list position -> original index), d_Xp[n*8] (list-ordered database).  Any output may be NULL
 * except d_N, d_O, d_Xp. */
int icp_rbc_construct(icp_ctx *ctx, const float *d_X, uint32_t n, const float *d_R, uint32_t nr, float alpha,
                      uint32_t *d_rep_id, uint32_t *d_N, uint32_t *d_O, uint32_t *d_perm, float *d_Xp);
/* RBC::RBCSearch<KINECT_R,GENERIC,KINECT>::run as wired at algorithms.cpp:4520-4536, :4674.
 * Outputs in rep-sorted query order: d_Qp[m*8], d_NN[m*8], d_NN_ID[m]; optional d_q_rep[m] (original order),
 * d_qperm[m] (sorted position -> original query), d_Nq[nr], d_Oq[nr]. */
int icp_rbc_search(icp_ctx *ctx, const float *d_Q, uint32_t m, const float *d_R, uint32_t nr, float alpha,
                   const float *d_Xp, const uint32_t *d_O, const uint32_t *d_N,
                   float *d_Qp, float *d_NN, icp_dist_id *d_NN_ID,
                   uint32_t *d_q_rep, uint32_t *d_qperm, uint32_t *d_Nq, uint32_t *d_Oq);
/* EXACT nearest neighbour over the random ball cover (the exact variant of the RBC paper, which the reference
 * leaves out -- README.md:4 "frame-to-frame" -- and frame-to-model mapping needs).  Same database as icp_rbc_search
 * (d_R, d_Xp, d_O, d_N from icp_rbc_construct; the database may be larger than the query set), but every list whose
 * representative cannot be excluded by the triangle inequality is searched, so the result is that of a brute-force scan of
 * X_p: smallest distance, lowest list position among ties.  Outputs in ORIGINAL query order: d_NN_ID[m] {dist, position in
 * X_p}, optional d_NN[m*8] (the matched points), optional d_evals[1] (+= distance evaluations executed, for the pruning rate). */
int icp_rbc_search_exact(icp_ctx *ctx, const float *d_Q, uint32_t m, const float *d_R, uint32_t nr, float alpha,
                         const float *d_Xp, const uint32_t *d_O, const uint32_t *d_N,
                         icp_dist_id *d_NN_ID, float *d_NN, uint64_t *d_evals);
/* ICPWeights::run, algorithms.hpp:485-572, kernels :212-254, :294-329 */
int icp_weights(icp_ctx *ctx, const icp_dist_id *d_in, float *d_W, double *d_sum_w, uint32_t n);
/* ICPMean<REGULAR>::run, algorithms.hpp:624-727, kernels :370-411, :529-566.  d_mean = 2 x float4 */
int icp_mean(icp_ctx *ctx, const float *d_F, const float *d_M, float *d_mean, uint32_t n);
/* ICPMean<WEIGHTED>::run, algorithms.hpp:733-837, kernel :454-495 */
int icp_mean_weighted(icp_ctx *ctx, const float *d_F, const float *d_M, const float *d_W, const double *d_sum_w,
                      float *d_mean, uint32_t n);
/* ICPDevs::run, algorithms.hpp:867-939, kernel :587-602.  d_DF, d_DM = n x float4 */
int icp_devs(icp_ctx *ctx, const float *d_F, const float *d_M, const float *d_mean, float *d_DF, float *d_DM, uint32_t n);
/* ICPS<REGULAR|WEIGHTED>::run, algorithms.hpp:990-1185, kernels :632-743 + reduce_sum_f.  d_W NULL => REGULAR */
int icp_sij(icp_ctx *ctx, const float *d_DM, const float *d_DF, const float *d_W, float *d_S11, uint32_t m, float c);
/* ICPPowerMethod::run, algorithms.hpp:1451-1537, kernel :976-1054 */
int icp_power_method(icp_ctx *ctx, const float *d_S11, const float *d_mean, float *d_Tk8);
/* rotation solve of ICPStep<EIGEN,*>::run (host Eigen code algorithms.cpp:3877-3896), on the device.
 * d_Rk9 (row-major) may be NULL. */
int icp_svd_solve(icp_ctx *ctx, const float *d_S11, const float *d_mean, float *d_Tk8, float *d_Rk9);
/* Reduce<MIN,float> / Reduce<MAX,uint> / Reduce<SUM,float>::run, algorithms.hpp:83-185, reduce_kernels.cl */
int icp_reduce_min_f(icp_ctx *ctx, const float *d_in, uint32_t cols, uint32_t rows, float *d_out);
int icp_reduce_max_ui(icp_ctx *ctx, const uint32_t *d_in, uint32_t cols, uint32_t rows, uint32_t *d_out);
int icp_reduce_sum_f(icp_ctx *ctx, const float *d_in, uint32_t cols, uint32_t rows, float *d_out);
/* Scan<INCLUSIVE|EXCLUSIVE,int>::run, algorithms.hpp:207-289, scan_kernels.cl */
int icp_scan_i(icp_ctx *ctx, const int32_t *d_in, uint32_t cols, uint32_t rows, int inclusive, int32_t *d_out);

/* ---- registration engine: ICPStep<CR,CW> (algorithms.hpp:1613-2401) and ICP<CR,CW> (:2433-2496) ---- */
enum { ICP_MEM_D_IN_F = 3, ICP_MEM_D_IN_M = 4, ICP_MEM_D_IO_T = 5 };   /* ICPStep::Memory, algorithms.hpp:2241-2267 */

int  icp_step_create(icp_ctx *ctx, int rot_cfg, int w_cfg, icp_step **out);
void icp_step_destroy(icp_step *s);
/* get(Memory) = buf before init(): share an external device buffer (algorithms.cpp:216-221 semantics) */
int  icp_step_bind(icp_step *s, int mem, void *d_ptr);
/* ICPStep::init (algorithms.cpp:4403-4582).  lm_w x lm_h = landmark grid used for the representative
 * sampling (0,0 => 128 x 128, the reference's fixed grid; lm_w*lm_h must equal m). */
int  icp_step_init(icp_step *s, uint32_t m, uint32_t nr, float alpha, float c, uint32_t lm_w, uint32_t lm_h);
void *icp_step_buffer(icp_step *s, int mem);                  /* device pointer behind get(Memory) */
/* write(D_IN_F|D_IN_M|D_IO_T, ptr) (algorithms.cpp:4596-4623): h2d from host memory.
 * As in the reference, writing D_IO_T only seeds the pose used by the next transform; the accumulated
 * R, t, s members are NOT updated from it (latent inconsistency of the reference, kept for parity). */
int  icp_step_write(icp_step *s, int mem, const void *h_src, int block);
int  icp_step_reset(icp_step *s);                             /* pose := identity, k := 0 (init() tail, :4485-4493) */
int  icp_step_set_alpha(icp_step *s, float alpha);            /* setAlpha, algorithms.cpp:4708-4713 */
int  icp_step_set_scaling(icp_step *s, float c);              /* setScaling, :4735-4739 */
int  icp_step_set_metric(icp_step *s, float f_g, float f_p);  /* override f_g(a), f_p(a) */
int  icp_step_set_mode(icp_step *s, int mode);                /* ICP_MODE_STAGED | ICP_MODE_FUSED */
int  icp_step_build_rbc(icp_step *s);                         /* buildRBC, algorithms.cpp:4655-4660 */
/* n_iters x ICPStep::run (algorithms.cpp:4670-4698) back to back on the device, pose accumulated on the
 * device, no host round trip.  Non-blocking. */
int  icp_step_run(icp_step *s, uint32_t n_iters);
/* ICP::run (algorithms.cpp:4807-4814): iterate until ICP::check() (:4824-4834) says stop, evaluated on the
 * device.  Blocking; *k_out = number of iterations performed (ICP::k). */
int  icp_run(icp_step *s, uint32_t max_iterations, double angle_threshold_deg, double translation_threshold_mm, uint32_t *k_out);
/* public state of ICPStep (algorithms.hpp:2302-2320): Rk[9] qk[4] tk[3] sk R[9] q[4] t[3] s (row-major 3x3),
 * then k.  Blocking read. */
typedef struct { float Rk[9], qk[4], tk[3], sk, R[9], q[4], t[3], s; uint32_t k; uint32_t done; } icp_state;
int  icp_step_get_state(icp_step *s, icp_state *h_out);
int  icp_step_get_pose_matrix(icp_step *s, float *h_T16);    /* [[s*R, t],[0 0 0 1]] row-major; blocking */
/* parity / profiling hooks: device pointers of the engine's intermediates, by name:
 * "reps" "rep_id" "N" "O" "perm" "Xp" "Mt" "q_rep" "qperm" "Nq" "Oq" "Qp" "NN" "NN_ID" "W" "sum_w" "mean"
 * "DF" "DM" "S" "Tk" "state".  NULL if unknown / not materialised in the current mode. */
void *icp_step_debug_ptr(icp_step *s, const char *name);
/* per-stage device time of the last staged run (ms), like run(timer): order = transform, rbc_search, weights,
 * means, devs, S, solve.  Runs ONE timed step (blocking). */
int  icp_step_run_timed(icp_step *s, float *h_ms7);
/* distance-evaluation counters (for the roofline): enable, run, read.  e1 = stage-1 evaluations (m*nr per
 * iteration), e2 = stage-2 evaluations (sum of searched list sizes), accumulated since the last reset. Blocking. */
int  icp_step_set_count_evals(icp_step *s, int on);
int  icp_step_eval_counts(icp_step *s, uint64_t *e1, uint64_t *e2);
/* stage-1 distance evaluations actually executed by the pruned kernel A (fused mode; <= e1), same accumulation. */
int  icp_step_stage1_executed(icp_step *s, uint64_t *e1x);
/* stage-2 distance evaluations actually executed (pruned walk of kernel A + list scans of kernel C; <= e2). */
int  icp_step_stage2_executed(icp_step *s, uint64_t *e2x);
/* measurement variants of icp_step_run: 0 = plain stream launches, 1 = unrolled CUDA graph,
 * 2 = conditional WHILE graph (device-side loop), 3 = persistent cooperative kernel (one launch, software grid barriers
 * between the phases; fused mode, chunk sizes <= 1024 points).  Same results. */
int  icp_step_run_variant(icp_step *s, uint32_t n_iters, int variant);

/* ---- batched registration: independent frame pairs, one engine slot per pair ---- */
int  icp_batch_create(icp_ctx *ctx, int rot_cfg, int w_cfg, uint32_t n_pairs, uint32_t m, uint32_t nr,
                      float alpha, float c, uint32_t lm_w, uint32_t lm_h, icp_batch **out);
void icp_batch_destroy(icp_batch *b);
float *icp_batch_F(icp_batch *b);                             /* device [n_pairs][m][8] */
float *icp_batch_M(icp_batch *b);
/* fill pair i with: F = d_base (m x 8), M = T_i^-1(F) + noise, T_i drawn from (seed + i) */
int  icp_batch_synthesize(icp_batch *b, const float *d_base, uint64_t seed);
int  icp_batch_upload(icp_batch *b, uint32_t first_pair, uint32_t count, const float *h_F, const float *h_M, int block);
/* buildRBC + n_iters steps for every pair; poses left on the device.  Non-blocking. */
int  icp_batch_register(icp_batch *b, uint32_t n_iters);
/* run icp_batch_register as n_slices concurrent slices of consecutive pairs (separate streams, fork/join on the
 * context stream): the latency-bound kernels of one slice overlap the FP32-bound kernels of the others.  Default 1. */
int  icp_batch_set_slices(icp_batch *b, uint32_t n_slices);
uint32_t icp_batch_slices(icp_batch *b);
/* kernel-C flavour the batch runs: 0 = k_search<L>, 1 = k_search_grouped, 2 = k_colscan_sort + k_search_sorted */
int  icp_batch_cmode(icp_batch *b);
/* host-buffer entry: h_F / h_M = [n_pairs][m][8] in host memory (pinned for full copy/compute overlap).  The batch
 * is cut into n_slices slices (0 = default); slice i+1 uploads on a copy stream while slice i registers.  Blocking;
 * h_T8 = [n_pairs][8] poses.  Same results as upload + register + read_poses. */
int  icp_batch_register_host(icp_batch *b, const float *h_F, const float *h_M, uint32_t n_iters, uint32_t n_slices, float *h_T8);
/* asynchronous pair of icp_batch_register_host (streaming callers): enqueue on the batch's own streams and return;
 * icp_batch_collect waits and copies the poses out.  Alternate two batches to overlap the uploads of one step with
 * the registration of the previous one.  Host buffers must stay valid until icp_batch_collect returns. */
int  icp_batch_register_host_async(icp_batch *b, const float *h_F, const float *h_M, uint32_t n_iters, uint32_t n_slices);
int  icp_batch_collect(icp_batch *b, float *h_T8 /*[n_pairs][8]*/);
int  icp_batch_read_poses(icp_batch *b, float *h_T8 /*[n_pairs][8]*/, float *h_T16 /*[n_pairs][16] or NULL*/);
void *icp_batch_debug_ptr(icp_batch *b, const char *name);
/* roofline hook: average device time (ms, CUDA events) of ONE fused kernel over the first n_launches iterations of a
 * fresh registration of the batch's pairs (the work per iteration depends on how far the registration has converged).
 * which: 0 = A assign (stage 1), 1 = B column scan / sort, 2 = C list search, 3 = D reduce + solve. */
int  icp_batch_time_kernel(icp_batch *b, int which, uint32_t n_launches, float *ms_avg);
int  icp_batch_config(icp_batch *b, uint32_t *QB, uint32_t *nbA, int *S, int *CL, int *L);

/* ---- single-process multi-GPU registration of independent frame pairs (replicas only; the reference has no
 *      multi-device code, its one-queue-per-object shape is algorithms.hpp:101 / :1650) ----
 * Device d owns a contiguous block of pairs end to end (its own icp_ctx + icp_batch, one host thread per device for the
 * duration of a call); no collective, only the 8-float poses come back.  Results do not depend on the device count. */
typedef struct icp_multi icp_multi;
/* n_devices = 0: every visible GPU; devices = NULL: ordinals 0 .. n_devices-1.  At least one pair per device. */
int  icp_multi_create(int n_devices, const int *devices, int rot_cfg, int w_cfg, uint32_t n_pairs, uint32_t m, uint32_t nr,
                      float alpha, float c, uint32_t lm_w, uint32_t lm_h, icp_multi **out);
void icp_multi_destroy(icp_multi *mg);
int  icp_multi_devices(icp_multi *mg);                        /* devices actually used */
int  icp_multi_pair_range(icp_multi *mg, int index, int *device, uint32_t *first, uint32_t *count);
/* h_F / h_M = [n_pairs][m][8] host memory (pinned for copy/compute overlap), h_T8 = [n_pairs][8] poses.  Blocking:
 * buildRBC + n_iters iterations of every pair, same results as icp_batch_register_host on one GPU. */
int  icp_multi_register_host(icp_multi *mg, const float *h_F, const float *h_M, uint32_t n_iters, float *h_T8);
/* one-shot convenience (create on n_devices GPUs, register, destroy); 128 x 128 landmark grid */
int  icp_multi_register_host_once(int n_devices, int rot_cfg, int w_cfg, uint32_t n_pairs, uint32_t m, uint32_t nr, float alpha, float c,
                                  const float *h_F, const float *h_M, uint32_t n_iters, float *h_T8);

/* micro-benchmark used for the FP32 roofline denominator: non-fused mul/add issue rate (flop/s). */
int  icp_measure_fp32_peak(icp_ctx *ctx, double *flops_scalar, double *flops_packed);
/* out4 = flop/s of {scalar mul+add, packed mul2+add2, scalar FFMA, packed FFMA2} */
int  icp_measure_fp32_rates(icp_ctx *ctx, double *out4);
/* launch/sync floor: average ms of an empty kernel launch chain and of a graph-replayed chain */
int  icp_measure_launch_floor(icp_ctx *ctx, float *us_stream_launch, float *us_graph_node);

#ifdef __cplusplus
}
#endif
#endif /* ICP_B200_H */
