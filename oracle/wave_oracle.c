/*
 * wave_oracle.c -- CPU oracle for the 2-D wavelet filterbank hot path.
 *
 * ==========================================================================================
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / `--impl reference` legs may build, load or call anything under oracle/, and
 * only as the checker / CPU baseline.  The product path (pytorch_wavelets_b200/) never imports
 * it and fails loudly when libb200wave.so (the CUDA library) is missing.
 * ==========================================================================================
 *
 * What it is: a plain-C restatement (fp32 and fp64) of the reference's per-level algorithms
 *   dwt/lowlevel.py   afb1d :91-172, sfb1d :226-271, AFB2D :336-347, SFB2D :671-680, mypad :28-88
 *   utils.py          reflect :146-163, symm_pad_1d :166-174
 *   dtcwt/lowlevel.py colfilter/rowfilter :70-94, coldfilt/rowdfilt :97-151,
 *                     colifilt/rowifilt :154-239, q2c :243-260, c2q :263-295
 *   dtcwt/transform_funcs.py fwd_j1 :98-121, inv_j1 :152-184, fwd_j2plus :226-249,
 *                     inv_j2plus :279-307, highs_to_orientations :61-72, orientations_to_highs :75-95
 *   scatternet/lowlevel.py ScatLayerj1_f.forward :76-111
 * written as explicit index arithmetic (no convolution library), one exported function per
 * C-ABI entry point of include/b200wave.h with the same argument meaning but HOST pointers.
 *
 * How it is pinned: tests/golden/ (.npz files) hold outputs of the reference itself (imported from
 * /root/reference in the build container by tests/golden/make_golden.py, with the pywt stand-in
 * of oracle/pywt_standin) for every entry point, all padding modes, odd sizes and the option
 * surface; tests/test_oracle_golden.py checks this oracle against them on CPU, and
 * tests/test_oracle_vs_reference.py checks it live against the reference when /root/reference
 * exists.  The reference's own tests pin against PyWavelets / the `dtcwt` numpy package,
 * neither of which is installed here (no network), so parity to *those* is via the reference.
 *
 * Build: oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off).  -ffp-contract=off matters: the
 * fused multiply-adds are written explicitly (fmaf/fma) in the tap order that reproduces the
 * reference CPU result bit-for-bit; the compiler must not invent or remove any.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum {
  ORC_MODE_ZERO = 0,
  ORC_MODE_SYMMETRIC = 1,
  ORC_MODE_PER = 2,
  ORC_MODE_CONSTANT = 3,
  ORC_MODE_REFLECT = 4,
  ORC_MODE_REPLICATE = 5,
  ORC_MODE_PERIODIC = 6
};
enum { ORC_EMODE = -1, ORC_ESIZE = -2, ORC_EARG = -3, ORC_EFILTER = -4, ORC_EINTERNAL = -7 };

#define SQRT2_D 1.4142135623730951

/* 0 (default): plane passes use the row-vectorised forms; 1: the plain *_line forms.  Both compute
 * identical values (tests/test_oracle_golden.py::test_line_and_row_forms_identical). */
static int orc_use_line_forms = 0;
void orc_set_line_forms(int on) { orc_use_line_forms = on; }

/* modes accepted by afb1d / sfb1d (dwt/lowlevel.py:134,155,165,263-264); anything else is the
 * reference's ValueError("Unkown pad type"). */
static int orc_mode_ok(int mode) {
  return mode == ORC_MODE_ZERO || mode == ORC_MODE_SYMMETRIC || mode == ORC_MODE_PER ||
         mode == ORC_MODE_REFLECT || mode == ORC_MODE_PERIODIC;
}

static long orc_floordiv(long a, long b) {
  long q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}

/* Boundary extension: index i of the extended signal -> index into [0,N) or -1 for "zero".
 *   symmetric : utils.py:146-163 reflect(i, -0.5, N-0.5) (half-sample, edge repeated, period 2N)
 *   reflect   : F.pad(mode='reflect') (whole-sample, period 2N-2), dwt/lowlevel.py:83-84
 *   periodic  : np.pad(mode='wrap'), dwt/lowlevel.py:62-81
 *   periodiz. : dwt/lowlevel.py:135-141 -- odd N first extended by repeating the last sample,
 *               then circular over the even length. */
static long orc_ext_index(long i, long N, int mode) {
  if (i >= 0 && i < N) return i;
  long p, r;
  switch (mode) {
    case ORC_MODE_SYMMETRIC:
      p = 2 * N; r = i % p; if (r < 0) r += p;
      return r < N ? r : p - 1 - r;
    case ORC_MODE_REFLECT:
      if (N == 1) return 0;
      p = 2 * N - 2; r = i % p; if (r < 0) r += p;
      return r < N ? r : p - r;
    case ORC_MODE_PERIODIC:
      r = i % N; if (r < 0) r += N;
      return r;
    case ORC_MODE_PER:
      p = N + (N & 1); r = i % p; if (r < 0) r += p;
      return r < N ? r : N - 1;
    default:
      return -1;
  }
}

/* pywt.dwt_coeff_len (third-party PyWavelets, unpinned `PyWavelets>=1.0.0` in the reference's
 * requirements.txt:3; published rule): ceil(N/2) for periodization, floor((N+L-1)/2) otherwise.
 * Call site: dwt/lowlevel.py:153. */
int orc_coeff_len(int n, int flen, int mode) {
  if (n < 1 || flen < 1) return ORC_ESIZE;
  return mode == ORC_MODE_PER ? (n + 1) / 2 : (n + flen - 1) / 2;
}
/* dwt/lowlevel.py:242-267: N = 2K (periodization) or 2K - L + 2 (conv_transpose2d, padding L-2). */
int orc_rec_len(int k, int flen, int mode) {
  if (k < 1 || flen < 1) return ORC_ESIZE;
  return mode == ORC_MODE_PER ? 2 * k : 2 * k - flen + 2;
}

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define T float
#define FMA(a, b, c) fmaf((a), (b), (c))
#define SQRT(x) sqrtf(x)
#define FN(name) CAT(name, _f32)
#include "wave_oracle_impl.h"
#undef T
#undef FMA
#undef SQRT
#undef FN

#define T double
#define FMA(a, b, c) fma((a), (b), (c))
#define SQRT(x) sqrt(x)
#define FN(name) CAT(name, _f64)
#include "wave_oracle_impl.h"
#undef T
#undef FMA
#undef SQRT
#undef FN
