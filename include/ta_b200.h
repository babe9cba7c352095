/* ta_b200.h -- C ABI of libta_b200.so, the B200 (sm_100a) backend for the
 * time-correlation hot path of MDAnalysis/transport-analysis.
 *
 * The reference has no FFI of its own (it is pure Python); the boundary this
 * library replaces is the MDAnalysis AnalysisBase template-method protocol as
 * implemented by the two reference classes, plus the one library function they
 * call.  Each entry point names the reference code it stands in for (paths
 * relative to the reference checkout):
 *
 *   ta_stage_begin            VelocityAutocorr._prepare      transport_analysis/velocityautocorr.py:142-153
 *                             ViscosityHelfand._prepare      transport_analysis/viscosity.py:111-142
 *   ta_stage_slot/_commit     VelocityAutocorr._single_frame transport_analysis/velocityautocorr.py:178-194
 *                             ViscosityHelfand._single_frame transport_analysis/viscosity.py:167-199
 *   ta_stage_bulk             the same per-frame copies, for readers that expose the whole
 *                             [frames, atoms, 3] array (MemoryReader recipe, transport_analysis/tests/utils.py:66-75)
 *   ta_vacf_fft               VelocityAutocorr._conclude_fft transport_analysis/velocityautocorr.py:208-215
 *                             + tidynamics.acf (un-vendored dependency, call site :211-213)
 *   ta_vacf_windowed          VelocityAutocorr._conclude_simple transport_analysis/velocityautocorr.py:217-238
 *   ta_helfand                ViscosityHelfand._conclude     transport_analysis/viscosity.py:201-233 (direct lag sums)
 *   ta_helfand_fft            the same, O(T log T): S1 - 2 S2 by FFT (the idea flagged as future work in
 *                             docs/tutorials/helfand_dev_toy_system.ipynb:134) with the reference's own sum
 *                             (viscosity.py:212-226) on every lag where that difference is not good to 1e-10
 *   ta_fetch_by_particle      results.vacf_by_particle / results.visc_by_particle
 *                             (velocityautocorr.py:145-147, viscosity.py:117-119, :229-231)
 *
 * Conventions
 *   - every function returns 0 (TA_OK) or a negative TA_ERR_* code; the text of
 *     the last failure is available from ta_last_error().  Nothing throws or
 *     exits across the ABI.
 *   - the caller owns every host pointer except the pinned slabs handed out by
 *     ta_stage_slot (library-owned, valid until the next ta_stage_slot /
 *     ta_stage_begin / ta_ctx_destroy).  The library owns all device memory,
 *     streams, events and NCCL communicators.
 *   - one context per analysis object; a context is not thread-safe.  All
 *     calls are synchronous on return except ta_stage_commit / ta_stage_bulk,
 *     which only enqueue (ta_stage_end or any compute call waits for them).
 *   - there is no CPU fallback: without a CUDA device ta_ctx_create fails.
 *   - particles ("atoms") are sharded over the context's devices in contiguous
 *     ranges; the only inter-device exchange is one NCCL all-reduce of the
 *     T-long atom-summed series (+ the atom count) per compute call.
 */
#ifndef TA_B200_H
#define TA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TA_OK 0
#define TA_ERR_INVALID (-1)
#define TA_ERR_CUDA (-2)
#define TA_ERR_NCCL (-3)
#define TA_ERR_UNSUPPORTED (-4)
#define TA_ERR_NOMEM (-5)

#define TA_DTYPE_F32 0
#define TA_DTYPE_F64 1

#define TA_PRECISION_FP64 0 /* reference parity: rel 1e-10 */
#define TA_PRECISION_FP32 1 /* optional: stated tolerance 1e-5 */

#define TA_LAYOUT_ATOM_MAJOR 0 /* out[natoms][T] */
#define TA_LAYOUT_LAG_MAJOR 1  /* out[T][natoms]  (the reference's array layout) */

#define TA_NCCL_ID_BYTES 128

typedef struct ta_ctx ta_ctx;

int ta_version(void);
int ta_device_count(int* count);
const char* ta_last_error(const ta_ctx* ctx); /* ctx may be NULL: error of the last failed create */

/* Single-process context over ndev local devices (ndev > 1 opens NCCL with ncclCommInitAll). */
int ta_ctx_create(int ndev, const int* devices, ta_ctx** out);
/* Multi-process mode (one process per GPU, e.g. under torchrun): rank 0 calls
 * ta_nccl_unique_id, the host program broadcasts the 128 bytes, every rank
 * calls ta_ctx_create_rank.  Each rank stages only ITS particles. */
int ta_nccl_unique_id(void* id128);
int ta_ctx_create_rank(int device, int rank, int nranks, const void* id128, ta_ctx** out);
void ta_ctx_destroy(ta_ctx* ctx);

/* Pin / unpin caller memory so ta_stage_bulk can DMA from it directly. */
int ta_host_register(void* ptr, uint64_t bytes);
int ta_host_unregister(void* ptr);

/* Declare the problem: T analysed frames, N particles (this process's share in
 * multi-process mode), D = len(dims) selected columns (dim_type), source dtype,
 * n_fields = 1 (velocities) or 2 (velocities + positions -> Helfand moment
 * g = (m v) x, needs masses[N]), arithmetic precision. */
int ta_stage_begin(ta_ctx* ctx, int64_t T, int64_t N, int D, const int* dims, int src_dtype,
                   int n_fields, const double* masses, int precision);
/* Next pinned slab, shaped [capacity][n_fields][N][3] of src_dtype; fill frames
 * 0..nframes-1 of it, then commit them as analysed frames frame0.. */
int ta_stage_slot(ta_ctx* ctx, void** host_ptr, int64_t* frames_capacity);
int ta_stage_commit(ta_ctx* ctx, int64_t frame0, int64_t nframes);
/* Whole-trajectory fast path: field f is a frame-major array [*][src_atoms][3]
 * of src_dtype; analysed frame i is source frame frame_first + i*frame_step,
 * particle j is source atom atom_first + j.  fields[1] is ignored when
 * n_fields == 1. */
int ta_stage_bulk(ta_ctx* ctx, const void* const* fields, int64_t src_atoms, int64_t atom_first,
                  int64_t frame_first, int64_t frame_step, int64_t nframes);
int ta_stage_end(ta_ctx* ctx); /* wait until every committed frame is resident in HBM */

/* Compute calls: ts_out[T] receives results.timeseries (atom mean over ALL
 * devices / ranks).  Per-particle results stay on the devices. */
int ta_vacf_fft(ta_ctx* ctx, double* ts_out);
int ta_vacf_windowed(ta_ctx* ctx, double* ts_out);
int ta_helfand(ta_ctx* ctx, const double* volumes /*[T]*/, double boltzmann, double temp_avg,
               double* ts_out);
/* O(T log T) route to the same Helfand MSD: sum (g[i]-g[i+k])^2 = S1[k] - 2 S2[k] with S2 from the FFT
 * autocorrelation kernel (idea noted in docs/tutorials/helfand_dev_toy_system.ipynb:134).  The difference
 * cancels (relative error ~ 30 eps sum g^2 / MSD[k]), so every lag whose MSD is below thr * sum g^2 is
 * re-evaluated with the exact sum of viscosity.py:212-226; when more than 2 % of a shard's lags need that the
 * direct kernel of ta_helfand does the shard.  Same 1e-10 bar as ta_helfand; FP64 only; TA_ERR_UNSUPPORTED
 * for T beyond ~29,000 (callers fall back to ta_helfand, which serves any T). */
int ta_helfand_fft(ta_ctx* ctx, const double* volumes /*[T]*/, double boltzmann, double temp_avg,
                   double* ts_out);
int ta_fetch_by_particle(ta_ctx* ctx, int64_t atom0, int64_t natoms, int layout, double* out);

/* Post-processing of the atom-mean timeseries of the LAST compute call, which stays on the device:
 * over the points w = start, start + step, ... < stop of (times[w], timeseries[w])
 *   *integral  trapezoid rule           self_diffusivity_gk before its / dim_fac   velocityautocorr.py:316-322
 *   running[n] running trapezoid integral, running[0] = initial (NULL: not wanted)  plot_running_integral :407-414
 *   *slope     least-squares slope of the timeseries over times                      the fit of viscosity.py:235-245
 * times[T] is a host array (the caller's self.times, or lag times for the fit).  Any of the three outputs may be NULL.
 * The Simpson variant (self_diffusivity_gk_odd, :354-360) stays on the host with scipy. */
int ta_green_kubo(ta_ctx* ctx, const double* times, int64_t start, int64_t stop, int64_t step, double initial,
                  double* integral, double* running, double* slope);

/* Device-side timing of whatever is enqueued between begin and end (CUDA
 * events on every shard's compute stream; ms = max over local devices). */
int ta_timer_begin(ta_ctx* ctx);
int ta_timer_end(ta_ctx* ctx, float* ms);
/* Device time of the kernels of the last compute call that do the correlation work -- K1, K2 or K3; K1 + K5 + K6 for
 * ta_helfand_fft -- from CUDA events recorded around those launches on their own stream (max over local devices). */
int ta_last_kernel_ms(ta_ctx* ctx, float* ms);
/* Probes for the roofline denominators bench.py reports beside MEASURED_PEAKS.json (which holds HBM and bf16 only):
 * ta_probe_fp64: FP64 FMA rate of the context's first device, TFLOP/s (independent DFMA chains on every SM, CUDA
 * events on the compute stream).  ta_probe_h2d: sustained rate of twelve back-to-back contiguous `bytes`-long
 * cudaMemcpyAsync from pinned host memory to the context's first device, GB/s (the ceiling the staging copies of
 * ta_stage_bulk are measured against; call it on all ranks at once to see the box's concurrent ceiling). */
int ta_probe_fp64(ta_ctx* ctx, double* tflops);
int ta_probe_h2d(ta_ctx* ctx, uint64_t bytes, double* gbps);
/* Overwrite a 512 MB scratch buffer so that nothing of the inputs stays in L2
 * (benchmark hygiene for workloads smaller than the 126 MB L2). */
int ta_flush_l2(ta_ctx* ctx);
/* Introspection for bench / tests: kernel launches issued by this context so
 * far, and the plan chosen for the FFT route (0s before the first call). */
int64_t ta_launch_count(const ta_ctx* ctx);
int ta_fft_plan_info(const ta_ctx* ctx, int* H, int* npasses, int* radices /*[12]*/,
                     int* threads, int* smem_bytes, int* grid);
/* 1 if the last FFT-route launch kept the per-thread streams of its output stage (parked residue-0 result, particle
 * sums, normalisation table) in tensor memory (FP64, 6,144 < T <= 10,240), else 0 */
int ta_k1_uses_tmem(const ta_ctx* ctx);
/* (particle, lag) pairs the last ta_helfand_fft evaluated with the exact sum (viscosity.py:212-226) because the
 * S1 - 2 S2 difference was not good to 1e-10 there; -1: the direct kernel took over the whole call. */
int64_t ta_helfand_fft_refined(const ta_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* TA_B200_H */
