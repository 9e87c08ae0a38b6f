/*
 * b200wave.h -- C ABI of libb200wave.so: the B200 (sm_100a) 2-D wavelet filterbank engine.
 *
 * This is the drop-in boundary for the ONE hot path of fbcotter/pytorch_wavelets:
 * the per-level separable analysis / synthesis filter banks behind DWTForward / DWTInverse,
 * DTCWTForward / DTCWTInverse and the ScatLayer magnitude epilogue.  The reference has no FFI
 * of its own; its narrowest stable interface is the torch.autograd.Function layer, so there is
 * exactly one entry point here per reference Function.forward (each cites the reference
 * file:line it replaces).  INTEGRATION.md shows the ctypes stub a maintainer of the reference
 * would add to bind them.
 *
 * Conventions
 *   - All image tensors are fp32, device memory, NCHW; "planes" = N*C independent images.
 *   - Filter taps are HOST pointers to fp32 arrays holding the taps exactly as the reference
 *     stores them in its module buffers (analysis filters time-reversed, synthesis filters
 *     as-is: reference dwt/lowlevel.py:916-920,970; dtcwt/lowlevel.py:58-67).  They are copied
 *     into kernel parameters (constant bank) by value; no device filter memory is needed.
 *   - The caller owns every buffer, including outputs; the library never allocates device
 *     memory, keeps no reference after return, has no global mutable state and is re-entrant.
 *   - Every call is asynchronous on the given CUDA stream (a cudaStream_t passed as void*;
 *     NULL = legacy default stream) and never synchronises the host.
 *   - Return value: 0 on success, a negative B200W_E* code otherwise (b200w_strerror()).
 *     Shape / mode validation mirrors the reference's Python exceptions; the Python shell
 *     raises the same exception types before calling, so C errors are defensive.
 *   - mode integers are the reference's mode_to_int() codes (dwt/lowlevel.py:274-290).
 */
#ifndef B200WAVE_H
#define B200WAVE_H

#ifdef __cplusplus
extern "C" {
#endif

#define B200W_VERSION 100 /* 0.1.0 */

/* reference dwt/lowlevel.py:274-290 */
enum {
  B200W_MODE_ZERO = 0,
  B200W_MODE_SYMMETRIC = 1,
  B200W_MODE_PERIODIZATION = 2,
  B200W_MODE_CONSTANT = 3, /* rejected by the filter banks, as in the reference */
  B200W_MODE_REFLECT = 4,
  B200W_MODE_REPLICATE = 5, /* rejected, as in the reference */
  B200W_MODE_PERIODIC = 6
};

enum {
  B200W_OK = 0,
  B200W_EMODE = -1,   /* unknown / unsupported padding mode  (reference: ValueError "Unkown pad type") */
  B200W_ESIZE = -2,   /* bad tensor size (reference: ValueError rows/cols multiple of 2 or 4)          */
  B200W_EARG = -3,    /* null pointer / inconsistent arguments                                        */
  B200W_EFILTER = -4, /* unsupported filter length                                                   */
  B200W_ECUDA = -5,   /* CUDA launch error (cudaGetLastError captured)                                */
  B200W_ENOTIMPL = -6 /* reference raises NotImplementedError here                                   */
};

#define B200W_MAX_TAPS 40 /* longest supported filter (db20) */

int b200w_version(void);
const char* b200w_strerror(int code);
/* last CUDA error string seen by this thread's most recent failing call ("" if none) */
const char* b200w_last_cuda_error(void);

/* pywt.dwt_coeff_len as used at reference dwt/lowlevel.py:153: ceil(n/2) for periodization,
 * floor((n+flen-1)/2) otherwise.  Returns <0 on bad input. */
int b200w_dwt_coeff_len(int n, int flen, int mode);
/* length produced by one synthesis level from k coefficients (reference dwt/lowlevel.py:242-267):
 * 2k for periodization else 2k - flen + 2. */
int b200w_dwt_rec_len(int k, int flen, int mode);

/* ---------------------------------------------------------------------------------------------
 * K1  one 2-D DWT analysis level.           Replaces AFB2D.forward, reference dwt/lowlevel.py:336-347
 *     (afb1d along W then along H, :91-172; boundary index generation mypad :28-88).
 *   x      (planes, H, W), row pitch x_pitch elements, plane stride x_plane_stride elements
 *   ll     (planes, Ho, Wo) with ll_pitch / ll_plane_stride (lets the caller keep padded internal levels)
 *   highs  (planes, 3, Ho, Wo) contiguous: [lh, hl, hh]; lh = low along W, high along H
 *   fw_*   taps applied along W (the module buffers named *_col -- reference quirk,
 *          dwt/transform2d.py:70-71 vs lowlevel.py:336), length Lw;  fh_* along H, length Lh.
 *   Ho = b200w_dwt_coeff_len(H, Lh, mode), Wo = b200w_dwt_coeff_len(W, Lw, mode).
 * Also computes SFB2D.backward (dwt/lowlevel.py:683-694) when given the synthesis taps.
 */
int b200w_dwt_afb2d(const float* x, long long x_plane_stride, int x_pitch,
                    float* ll, long long ll_plane_stride, int ll_pitch,
                    float* highs,
                    int planes, int H, int W,
                    const float* fw_lo, const float* fw_hi, int Lw,
                    const float* fh_lo, const float* fh_hi, int Lh,
                    int mode, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1xJ  all J analysis levels in one call.  Replaces the level loop of DWTForward.forward, reference
 *     dwt/transform2d.py:68-74 (J x AFB2D.apply with the low-pass fed back).
 *   x      (planes, H, W) pitched as in K1
 *   yl     (planes, H_J, W_J) contiguous: the final low-pass
 *   highs  HOST array of J device pointers; highs[j] is level j+1's (planes, 3, H_j, W_j) contiguous tensor
 *          (finest first, the reference's yh list); H_j = b200w_dwt_coeff_len(H_{j-1}, Lh, mode) etc.
 *   When the fused pyramid kernel applies (equal even filter lengths <= 16, mode zero / symmetric / reflect,
 *   16-byte aligned rows with W % 4 == 0, every level at least as large as the filter, the plan fits shared
 *   memory) this is ONE kernel launch and the inter-level low-passes never touch device memory.  Otherwise the
 *   levels run one K1 launch each and their intermediate low-passes live in `workspace` (caller-owned device
 *   memory, at least b200w_dwt_forward_workspace(...) bytes, which is 0 when the fused kernel applies; the
 *   same x / strides must be passed to both calls since the answer depends on the alignment of x).
 */
long long b200w_dwt_forward_workspace(const float* x, long long x_plane_stride, int x_pitch,
                                      int planes, int H, int W, int J, int Lw, int Lh, int mode);
int b200w_dwt_forward(const float* x, long long x_plane_stride, int x_pitch,
                      int planes, int H, int W, int J,
                      float* yl, float* const* highs,
                      const float* fw_lo, const float* fw_hi, int Lw,
                      const float* fh_lo, const float* fh_hi, int Lh,
                      int mode, void* workspace, long long workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2  one 2-D DWT synthesis level.          Replaces SFB2D.forward, reference dwt/lowlevel.py:671-680
 *     (sfb1d along H for (ll,lh) and (hl,hh), then along W; :226-271).
 *   ll     (planes, Hc, Wc) pitched;  highs (planes, 3, Hc, Wc) contiguous, or NULL = zeros
 *          (reference dwt/transform2d.py:137-139)
 *   y      (planes, Ho, Wo) pitched.  Ho/Wo may be smaller than the natural size
 *          b200w_dwt_rec_len(Hc, Lh, mode) -- the crop of AFB2D.backward (lowlevel.py:359-364).
 *   gh_* taps along H (first pass), gw_* along W.
 */
int b200w_dwt_sfb2d(const float* ll, long long ll_plane_stride, int ll_pitch,
                    const float* highs,
                    float* y, long long y_plane_stride, int y_pitch,
                    int planes, int Hc, int Wc, int Ho, int Wo,
                    const float* gh_lo, const float* gh_hi, int Lh,
                    const float* gw_lo, const float* gw_hi, int Lw,
                    int mode, void* stream);

/* ---------------------------------------------------------------------------------------------
 * 1-D DWT levels.                           Replace AFB1D.forward / SFB1D.forward, reference dwt/lowlevel.py:388-404,
 *     717-729 (afb1d / sfb1d along the last dimension; the level loops are DWT1DForward / DWT1DInverse,
 *     dwt/transform1d.py:44-65, 97-115).  Same modes, tap conventions and length rules as K1 / K2.
 *   b200w_dwt_afb1d: x (rows, N) with row pitch x_pitch -> lo, hi (rows, K) contiguous, K = b200w_dwt_coeff_len(N, L, mode)
 *   b200w_dwt_sfb1d: lo, hi (rows, K) contiguous (hi may be NULL = zeros) -> y (rows, Nout) contiguous,
 *                    Nout <= b200w_dwt_rec_len(K, L, mode) (smaller = the crop of AFB1D.backward, :406-424)
 * Each also computes the other's backward pass when given the same stored filters.
 */
int b200w_dwt_afb1d(const float* x, long long x_pitch, int rows, int N, float* lo, float* hi,
                    const float* f0, const float* f1, int L, int mode, void* stream);
int b200w_dwt_sfb1d(const float* lo, const float* hi, int rows, int K, float* y, int Nout,
                    const float* g0, const float* g1, int L, int mode, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DTCWT.  "highs" is the reference's 6-D complex band-pass tensor; because o_dim / ri_dim are
 * configurable (dtcwt/transform_funcs.py:10-58) it is described by six ELEMENT strides
 * hs[6] = {n, c, orientation, row, col, real/imag}.  Default layout (N,C,6,H/2,W/2,2) is the
 * fast path.  Orientation order 15,45,75,105,135,165 degrees (transform_funcs.py:61-72).
 *
 * K3  level-1 forward.                     Replaces FWD_J1.forward, dtcwt/transform_funcs.py:346-358
 *     (fwd_j1 :98-121 = rowfilter x2, colfilter x4 (dtcwt/lowlevel.py:70-94), q2c :243-260).
 *   x   (N*C, H, W) pitched, H and W even;  ll (N*C, H, W) pitched
 *   highs: band-pass output at (H/2, W/2) or NULL when skip_hps.
 *   h0,h1: stored (reversed) level-1 filters, odd lengths L0, L1.
 *   mode: B200W_MODE_SYMMETRIC -> symmetric extension, anything else -> zero padding
 *         (dtcwt/lowlevel.py:75-79).
 * Also INV_J1.backward (:434-449) when given g0o/g1o.
 */
int b200w_dtcwt_fwd_j1(const float* x, long long x_plane_stride, int x_pitch,
                       float* ll, long long ll_plane_stride, int ll_pitch,
                       float* highs, const long long hs[6],
                       int N, int C, int H, int W,
                       const float* h0, int L0, const float* h1, int L1,
                       int mode, void* stream);

/* K4  level>=2 forward.                    Replaces FWD_J2PLUS.forward, transform_funcs.py:380-392
 *     (fwd_j2plus :226-249 = rowdfilt x2, coldfilt x4 (dtcwt/lowlevel.py:97-151), q2c).
 *   x (N*C,H,W) with H%4==0 and W%4==0 (else B200W_ESIZE, reference ValueError lowlevel.py:102-104);
 *   ll (N*C,H/2,W/2); highs at (H/4,W/4) or NULL when skip_hps.
 *   h0a,h1a,h0b,h1b: stored (reversed) q-shift filters, common even length m.
 *   Always symmetric extension (transform_funcs.py:381).
 * Also INV_J2PLUS.backward (:471-488) with a<->b swapped by the caller.
 */
int b200w_dtcwt_fwd_j2plus(const float* x, long long x_plane_stride, int x_pitch,
                           float* ll, long long ll_plane_stride, int ll_pitch,
                           float* highs, const long long hs[6],
                           int N, int C, int H, int W,
                           const float* h0a, const float* h1a,
                           const float* h0b, const float* h1b, int m,
                           void* stream);

/* K5  level-1 inverse.                     Replaces INV_J1.forward, transform_funcs.py:419-431
 *     (inv_j1 :152-184 = c2q (dtcwt/lowlevel.py:263-295), colfilter x4, rowfilter x2).
 *   ll (N*C,H,W) pitched or NULL (treated as zeros); highs at (H/2,W/2) or NULL (low-pass only
 *   path, which the reference runs with symmetric extension regardless of mode, :159);
 *   y (N*C,H,W).  g0,g1 stored level-1 synthesis filters (odd lengths).
 * Also FWD_J1.backward (:361-374) when given h0o/h1o.
 */
int b200w_dtcwt_inv_j1(const float* ll, long long ll_plane_stride, int ll_pitch,
                       const float* highs, const long long hs[6],
                       float* y, long long y_plane_stride, int y_pitch,
                       int N, int C, int H, int W,
                       const float* g0, int L0, const float* g1, int L1,
                       int mode, void* stream);

/* K6  level>=2 inverse.                    Replaces INV_J2PLUS.forward, transform_funcs.py:455-468
 *     (inv_j2plus :279-307 = c2q, colifilt x4, rowifilt x2 (dtcwt/lowlevel.py:154-239)).
 *   ll (N*C,H,W) or NULL; highs at (H/2,W/2) or NULL; y (N*C,2H,2W).  H, W even.
 *   g0a,g1a,g0b,g1b stored q-shift synthesis filters, common even length m.
 * Also FWD_J2PLUS.backward (:395-413) with a<->b swapped by the caller.
 */
int b200w_dtcwt_inv_j2plus(const float* ll, long long ll_plane_stride, int ll_pitch,
                           const float* highs, const long long hs[6],
                           float* y, long long y_plane_stride, int y_pitch,
                           int N, int C, int H, int W,
                           const float* g0a, const float* g1a,
                           const float* g0b, const float* g1b, int m,
                           void* stream);

/* K7  ScatLayer forward.                   Replaces ScatLayerj1_f.forward (combine_colour=False),
 *     scatternet/lowlevel.py:76-111: level-1 DTCWT (o_dim=1) -> 2x2 mean of ll ->
 *     sqrt(re^2+im^2+b^2)-b -> stacked (N,7,C,H/2,W/2).
 *   x (N,C,H,W) contiguous, H and W even; z (N,7,C,H/2,W/2) contiguous.
 *   dre_dr / dim_dr: optional (N,6,C,H/2,W/2) outputs re/r and im/r saved for the backward
 *   pass (:96-99); pass NULL when no gradient is needed.
 */
int b200w_scat_j1(const float* x, float* z, float* dre_dr, float* dim_dr,
                  int N, int C, int H, int W,
                  const float* h0, int L0, const float* h1, int L1,
                  int mode, float magbias, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU.  Every (n, c) plane is transformed independently (depthwise filters, no halo between planes:
 * reference dwt/lowlevel.py:143,164,168; dtcwt/lowlevel.py:77,111), so the batch shards over the GPUs of a
 * box with NO exchange during compute; the one collective of the path is the all-gather of the returned
 * tensors along dim 0 after the last level.  One process per GPU; the communicator is an explicit handle
 * (no hidden global state).  NCCL is bound at run time; without it these return B200W_ENOTIMPL.
 *   b200w_comm_unique_id  rank 0 creates the 128-byte NCCL id and hands it to the other ranks out of band
 *                         (the Python shell broadcasts it through torch.distributed)
 *   b200w_comm_init       collective over all ranks; binds the communicator to the CURRENT CUDA device
 *   b200w_allgather       recv[r*count .. (r+1)*count) = rank r's send[0 .. count)  (fp32 elements), asynchronous
 *                         on `stream` (pass the compute stream: the gather then simply follows the last level)
 */
typedef struct b200w_comm b200w_comm;
int b200w_comm_unique_id(void* id128);
int b200w_comm_init(b200w_comm** comm, int rank, int world, const void* id128);
int b200w_comm_destroy(b200w_comm* comm);
int b200w_allgather(b200w_comm* comm, const float* send, float* recv, long long count, void* stream);
const char* b200w_comm_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Generic-kernel variants.  Every entry point above picks a specialised streaming kernel when one
 * exists for the filter lengths / layout / alignment it is given and otherwise runs the generic tile
 * kernel (any filter length <= B200W_MAX_TAPS, any mode, any layout).  The *_generic symbols take the
 * same arguments and always run the generic tile kernel: an independent second implementation of the
 * same arithmetic, used by the parity tests for A/B checks (selection is per call -- the library keeps
 * no mode switch or any other global mutable state).
 */
int b200w_dwt_afb2d_generic(const float* x, long long x_plane_stride, int x_pitch,
                            float* ll, long long ll_plane_stride, int ll_pitch, float* highs,
                            int planes, int H, int W,
                            const float* fw_lo, const float* fw_hi, int Lw,
                            const float* fh_lo, const float* fh_hi, int Lh, int mode, void* stream);
int b200w_dwt_forward_generic(const float* x, long long x_plane_stride, int x_pitch,
                              int planes, int H, int W, int J, float* yl, float* const* highs,
                              const float* fw_lo, const float* fw_hi, int Lw,
                              const float* fh_lo, const float* fh_hi, int Lh,
                              int mode, void* workspace, long long workspace_bytes, void* stream);
int b200w_dwt_sfb2d_generic(const float* ll, long long ll_plane_stride, int ll_pitch, const float* highs,
                            float* y, long long y_plane_stride, int y_pitch,
                            int planes, int Hc, int Wc, int Ho, int Wo,
                            const float* gh_lo, const float* gh_hi, int Lh,
                            const float* gw_lo, const float* gw_hi, int Lw, int mode, void* stream);
int b200w_dtcwt_fwd_j1_generic(const float* x, long long x_plane_stride, int x_pitch,
                               float* ll, long long ll_plane_stride, int ll_pitch,
                               float* highs, const long long hs[6], int N, int C, int H, int W,
                               const float* h0, int L0, const float* h1, int L1, int mode, void* stream);
int b200w_dtcwt_fwd_j2plus_generic(const float* x, long long x_plane_stride, int x_pitch,
                                   float* ll, long long ll_plane_stride, int ll_pitch,
                                   float* highs, const long long hs[6], int N, int C, int H, int W,
                                   const float* h0a, const float* h1a, const float* h0b, const float* h1b,
                                   int m, void* stream);
int b200w_dtcwt_inv_j1_generic(const float* ll, long long ll_plane_stride, int ll_pitch,
                               const float* highs, const long long hs[6],
                               float* y, long long y_plane_stride, int y_pitch, int N, int C, int H, int W,
                               const float* g0, int L0, const float* g1, int L1, int mode, void* stream);
int b200w_dtcwt_inv_j2plus_generic(const float* ll, long long ll_plane_stride, int ll_pitch,
                                   const float* highs, const long long hs[6],
                                   float* y, long long y_plane_stride, int y_pitch, int N, int C, int H, int W,
                                   const float* g0a, const float* g1a, const float* g0b, const float* g1b,
                                   int m, void* stream);
int b200w_scat_j1_generic(const float* x, float* z, float* dre_dr, float* dim_dr, int N, int C, int H, int W,
                          const float* h0, int L0, const float* h1, int L1, int mode, float magbias, void* stream);

/* ---------------------------------------------------------------------------------------------
 * float64 variants.  The reference computes in torch's default dtype (dwt/lowlevel.py:972,
 * dtcwt/lowlevel.py:67), so double-precision modules work there.  Here double precision runs the
 * generic tile kernels (and the 1-D row kernels) compiled for `double`: same arguments as the
 * float32 entry points with `double` in place of `float`, same return codes, same stream semantics.
 * (The streaming / pyramid fast paths and b200w_dwt_forward are float32-only.)
 */
int b200w_dwt_afb2d_f64(const double* x, long long x_plane_stride, int x_pitch,
                        double* ll, long long ll_plane_stride, int ll_pitch, double* highs,
                        int planes, int H, int W,
                        const double* fw_lo, const double* fw_hi, int Lw,
                        const double* fh_lo, const double* fh_hi, int Lh, int mode, void* stream);
int b200w_dwt_sfb2d_f64(const double* ll, long long ll_plane_stride, int ll_pitch, const double* highs,
                        double* y, long long y_plane_stride, int y_pitch,
                        int planes, int Hc, int Wc, int Ho, int Wo,
                        const double* gh_lo, const double* gh_hi, int Lh,
                        const double* gw_lo, const double* gw_hi, int Lw, int mode, void* stream);
int b200w_dwt_afb1d_f64(const double* x, long long x_pitch, int rows, int N, double* lo, double* hi,
                        const double* f0, const double* f1, int L, int mode, void* stream);
int b200w_dwt_sfb1d_f64(const double* lo, const double* hi, int rows, int K, double* y, int Nout,
                        const double* g0, const double* g1, int L, int mode, void* stream);
int b200w_dtcwt_fwd_j1_f64(const double* x, long long x_plane_stride, int x_pitch,
                           double* ll, long long ll_plane_stride, int ll_pitch,
                           double* highs, const long long hs[6], int N, int C, int H, int W,
                           const double* h0, int L0, const double* h1, int L1, int mode, void* stream);
int b200w_dtcwt_fwd_j2plus_f64(const double* x, long long x_plane_stride, int x_pitch,
                               double* ll, long long ll_plane_stride, int ll_pitch,
                               double* highs, const long long hs[6], int N, int C, int H, int W,
                               const double* h0a, const double* h1a, const double* h0b, const double* h1b,
                               int m, void* stream);
int b200w_dtcwt_inv_j1_f64(const double* ll, long long ll_plane_stride, int ll_pitch,
                           const double* highs, const long long hs[6],
                           double* y, long long y_plane_stride, int y_pitch, int N, int C, int H, int W,
                           const double* g0, int L0, const double* g1, int L1, int mode, void* stream);
int b200w_dtcwt_inv_j2plus_f64(const double* ll, long long ll_plane_stride, int ll_pitch,
                               const double* highs, const long long hs[6],
                               double* y, long long y_plane_stride, int y_pitch, int N, int C, int H, int W,
                               const double* g0a, const double* g1a, const double* g0b, const double* g1b,
                               int m, void* stream);
int b200w_scat_j1_f64(const double* x, double* z, double* dre_dr, double* dim_dr, int N, int C, int H, int W,
                      const double* h0, int L0, const double* h1, int L1, int mode, double magbias, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Standalone 1-D DTCWT primitives (the reference's low-level API; in the transforms these passes are
 * fused inside the per-level kernels).  x: (planes, H, W) contiguous; y: contiguous output.
 *   b200w_dtcwt_filter  = colfilter (along_w = 0) / rowfilter (along_w = 1), reference
 *                         dtcwt/lowlevel.py:70-94: y[n] = sum_j h[j] x[ext(n + j - L/2)], ext = symmetric
 *                         or zero padding; ANY length L -- an even L yields N + 1 outputs along the
 *                         filtered dimension (the reference's behaviour, tests/test_colfilter.py:52-61).
 *   b200w_dtcwt_dfilt   = coldfilt / rowdfilt (:97-151): (ha, hb) even length m, filtered size % 4 == 0,
 *                         output half size.
 *   b200w_dtcwt_ifilt   = colifilt / rowifilt (:154-239): filtered size % 2 == 0, output double size.
 * Taps are host pointers in stored (reversed) order, as everywhere in this ABI.
 */
int b200w_dtcwt_filter(const float* x, float* y, int planes, int H, int W, const float* h, int L,
                       int symmetric, int along_w, void* stream);
int b200w_dtcwt_dfilt(const float* x, float* y, int planes, int H, int W, const float* ha, const float* hb,
                      int m, int highpass, int along_w, void* stream);
int b200w_dtcwt_ifilt(const float* x, float* y, int planes, int H, int W, const float* ha, const float* hb,
                      int m, int highpass, int along_w, void* stream);
int b200w_dtcwt_filter_f64(const double* x, double* y, int planes, int H, int W, const double* h, int L,
                           int symmetric, int along_w, void* stream);
int b200w_dtcwt_dfilt_f64(const double* x, double* y, int planes, int H, int W, const double* ha,
                          const double* hb, int m, int highpass, int along_w, void* stream);
int b200w_dtcwt_ifilt_f64(const double* x, double* y, int planes, int H, int W, const double* ha,
                          const double* hb, int m, int highpass, int along_w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200WAVE_H */
