/*
 * wave_oracle_impl.h -- type-generic body of the CPU oracle (included twice by wave_oracle.c,
 * once with T=float and once with T=double).
 *
 * TEST INFRASTRUCTURE ONLY.  See wave_oracle.c for the header and the rules about who may use it.
 * Every function restates the reference algorithm and cites the reference file:line it follows
 * (paths relative to /root/reference/pytorch_wavelets/).
 *
 * Required macros: T, FMA(a,b,c), SQRT(x), FN(name) (adds the type suffix).
 */

/* ---- 1-D building blocks ------------------------------------------------------------------- */

/* dwt/lowlevel.py:91-172 afb1d, one line.  f = stored (time-reversed) taps, correlation form:
 *   out[k] = sum_j f[j] * xe[2k + j - pl],  pl = L-2 (non-periodization, :153-168: pad p//2 with
 *   p = 2(K-1)-N+L) or L-1-L//2 (periodization, :134-150: roll by -L//2, pad L-1, fold).
 * Accumulation order = increasing stored index j (increasing input index), first term a plain
 * product, then fused multiply-adds: the order that reproduces the reference CPU result
 * bit-for-bit. */
static void FN(afb_line)(const T* x, long xs, int N, T* lo, long los, T* hi, long his, int K,
                         const T* f0, const T* f1, int L, int mode) {
  const int pl = (mode == ORC_MODE_PER) ? (L - 1 - L / 2) : (L - 2);
  for (int k = 0; k < K; ++k) {
    T a0 = 0, a1 = 0;
    for (int j = 0; j < L; ++j) {
      long i = orc_ext_index(2L * k + j - pl, N, mode);
      T v = (i < 0) ? (T)0 : x[i * xs];
      if (j == 0) { a0 = f0[0] * v; a1 = f1[0] * v; }
      else { a0 = FMA(f0[j], v, a0); a1 = FMA(f1[j], v, a1); }
    }
    lo[k * los] = a0;
    hi[k * his] = a1;
  }
}

/* dwt/lowlevel.py:226-271 sfb1d, one line.  g = stored (un-reversed) taps.
 *   y[n] = sum_k lo[k] g0[s-2k] + sum_k hi[k] g1[s-2k],  s = n + off,
 *   off = L-2 (conv_transpose2d padding L-2, :263-267) or L/2-1 (periodization: fold + roll, :252-261),
 *   k outside [0,K) contributes nothing (non-per) or wraps mod K (per).
 * Order (defined here; the reference leaves it to oneDNN): each of the two transposed convolutions is
 * accumulated over increasing k, then the two are added -- mirroring "conv_transpose(lo)+conv_transpose(hi)". */
static void FN(sfb_line)(const T* lo, long los, const T* hi, long his, int K, T* y, long ys, int Nout,
                         const T* g0, const T* g1, int L, int mode) {
  const int off = (mode == ORC_MODE_PER) ? (L / 2 - 1) : (L - 2);
  for (int n = 0; n < Nout; ++n) {
    long s = (long)n + off;
    /* t = s - 2k in [0, L)  <=>  k in [ceil((s-L+1)/2), floor(s/2)] */
    long kmin = orc_floordiv(s - L + 2, 2);
    long kmax = orc_floordiv(s, 2);
    T a0 = 0, a1 = 0;
    int first = 1;
    for (long k = kmin; k <= kmax; ++k) {
      long kk = k;
      if (mode == ORC_MODE_PER) { kk = k % K; if (kk < 0) kk += K; }
      else if (k < 0 || k >= K) continue;
      int t = (int)(s - 2 * k);
      T vl = lo ? lo[kk * los] : (T)0;
      T vh = hi ? hi[kk * his] : (T)0;
      if (first) { a0 = vl * g0[t]; a1 = vh * g1[t]; first = 0; }
      else { a0 = FMA(vl, g0[t], a0); a1 = FMA(vh, g1[t], a1); }
    }
    y[n * ys] = a0 + a1;
  }
}

/* dtcwt/lowlevel.py:70-94 colfilter/rowfilter, one line.  h = stored (reversed) taps, m = L//2.
 *   symmetric (:75-77): y[n] = sum_j h[j] x[sym(n + j - m)],  n in [0, N + 2m - L + 1)
 *   otherwise (:78-79): zero padding m each side.
 * accumulate != 0: y[n] = y[n] + result (the "+" of transform_funcs.py:166-182, second operand). */
static void FN(filt_line)(const T* x, long xs, int N, T* y, long ys, const T* h, int L, int symmetric,
                          int accumulate_first) {
  const int m = L / 2;
  const int Nout = N + 2 * m - L + 1;
  for (int n = 0; n < Nout; ++n) {
    T a = 0;
    for (int j = 0; j < L; ++j) {
      long i = (long)n + j - m;
      if (symmetric) i = orc_ext_index(i, N, ORC_MODE_SYMMETRIC);
      else if (i < 0 || i >= N) i = -1;
      T v = (i < 0) ? (T)0 : x[i * xs];
      a = (j == 0) ? h[0] * v : FMA(h[j], v, a);
    }
    /* accumulate_first: existing y is the FIRST operand of the reference's sum (y_prev + a) */
    y[n * ys] = accumulate_first ? (y[n * ys] + a) : a;
  }
}

/* dtcwt/lowlevel.py:97-151 coldfilt/rowdfilt, one line; (ha, hb) in the callee's argument order,
 * stored (reversed), common even length m; N % 4 == 0.
 *   Ya[q] = sum_j ha[j] x[sym(4q + 2j + 2 - m)]   (xe[2::2], :109)
 *   Yb[q] = sum_j hb[j] x[sym(4q + 2j + 3 - m)]   (xe[3::2])
 *   lowpass: y[2q]=Ya, y[2q+1]=Yb;  highpass: y[2q]=Yb, y[2q+1]=Ya  (:117-120) */
static void FN(dfilt_line)(const T* x, long xs, int N, T* y, long ys, const T* ha, const T* hb,
                           int m, int highpass) {
  for (int q = 0; q < N / 4; ++q) {
    T a = 0, b = 0;
    for (int j = 0; j < m; ++j) {
      long ia = orc_ext_index(4L * q + 2 * j + 2 - m, N, ORC_MODE_SYMMETRIC);
      long ib = orc_ext_index(4L * q + 2 * j + 3 - m, N, ORC_MODE_SYMMETRIC);
      T va = x[ia * xs], vb = x[ib * xs];
      if (j == 0) { a = ha[0] * va; b = hb[0] * vb; }
      else { a = FMA(ha[j], va, a); b = FMA(hb[j], vb, b); }
    }
    if (highpass) { y[(2L * q) * ys] = b; y[(2L * q + 1) * ys] = a; }
    else { y[(2L * q) * ys] = a; y[(2L * q + 1) * ys] = b; }
  }
}

/* dtcwt/lowlevel.py:154-239 colifilt/rowifilt, one line; (ha, hb) callee order, stored, even length m,
 * m2 = m/2, N even, output 2N.  xe[i] = sym(i - m2) (:167).
 *   y[4t+s] = sum_{j<m2} f_s[j] x[xe[2(t+j) + o_s]]
 *   m2 even (:169-177): f = (hae, hbe, hao, hbo), o = (0,1,2,3) lowpass / (1,0,3,2) highpass
 *   m2 odd  (:178-186): f = (hao, hbo, hae, hbe), o = (1,2,1,2) lowpass / (2,1,2,1) highpass
 *   with hae[j] = ha[2j], hao[j] = ha[2j+1] on the stored arrays (:159-162). */
static void FN(ifilt_line)(const T* x, long xs, int N, T* y, long ys, const T* ha, const T* hb,
                           int m, int highpass, int accumulate_first) {
  const int m2 = m / 2;
  int o[4], par[4]; /* par: 0 -> even taps (h?e), 1 -> odd taps (h?o) */
  if (m2 % 2 == 0) {
    par[0] = 0; par[1] = 0; par[2] = 1; par[3] = 1;
    if (highpass) { o[0] = 1; o[1] = 0; o[2] = 3; o[3] = 2; } else { o[0] = 0; o[1] = 1; o[2] = 2; o[3] = 3; }
  } else {
    par[0] = 1; par[1] = 1; par[2] = 0; par[3] = 0;
    if (highpass) { o[0] = 2; o[1] = 1; o[2] = 2; o[3] = 1; } else { o[0] = 1; o[1] = 2; o[2] = 1; o[3] = 2; }
  }
  for (int t = 0; t < N / 2; ++t) {
    for (int s = 0; s < 4; ++s) {
      const T* h = (s & 1) ? hb : ha;
      T a = 0;
      for (int j = 0; j < m2; ++j) {
        long i = orc_ext_index(2L * (t + j) + o[s] - m2, N, ORC_MODE_SYMMETRIC);
        T v = x[i * xs];
        T c = h[2 * j + par[s]];
        a = (j == 0) ? c * v : FMA(c, v, a);
      }
      long yi = (4L * t + s) * ys;
      y[yi] = accumulate_first ? (y[yi] + a) : a;
    }
  }
}

/* ---- row-vectorised forms of the same arithmetic -------------------------------------------------
 * The *_line functions above are the readable statement of each 1-D operator.  The plane passes
 * below compute exactly the same values (same products, same fused multiply-adds, same order over
 * the tap index) but walk memory row-wise so the compiler can vectorise across independent outputs:
 *   - along W: extend one row into a scratch line once, then taps outer / outputs inner;
 *   - along H: one output row at a time, taps outer / columns inner.
 * This is what makes the oracle usable as the CPU baseline (bench.py) without flattering the GPU. */

/* build e[i] = x[ext(i + start)] (or 0) for i in [0, n) */
static void FN(extend_line)(const T* x, int N, T* e, long start, int n, int mode) {
  for (int i = 0; i < n; ++i) {
    long g = orc_ext_index(start + i, N, mode);
    e[i] = (g < 0) ? (T)0 : x[g];
  }
}

/* afb along W for all rows of a plane */
static void FN(afb_rows)(const T* x, long xpitch, int H, int W, T* lo, T* hi, int Wo, const T* f0, const T* f1,
                         int L, int mode, T* e) {
  const int pl = (mode == ORC_MODE_PER) ? (L - 1 - L / 2) : (L - 2);
  const int n = 2 * (Wo - 1) + L;
  for (int r = 0; r < H; ++r) {
    FN(extend_line)(x + (long)r * xpitch, W, e, -pl, n, mode);
    T* l = lo + (long)r * Wo;
    T* h = hi + (long)r * Wo;
    for (int k = 0; k < Wo; ++k) { T v = e[2 * k]; l[k] = f0[0] * v; h[k] = f1[0] * v; }
    for (int j = 1; j < L; ++j) {
      const T c0 = f0[j], c1 = f1[j];
      for (int k = 0; k < Wo; ++k) { T v = e[2 * k + j]; l[k] = FMA(c0, v, l[k]); h[k] = FMA(c1, v, h[k]); }
    }
  }
}

/* afb along H: in (H, Wc) pitch inpitch -> a (Ho, Wc) pitch ap, b (Ho, Wc) pitch bp */
static void FN(afb_cols)(const T* in, long inpitch, int H, int Wc, T* a, long ap, T* b, long bp, int Ho,
                         const T* f0, const T* f1, int L, int mode) {
  const int pl = (mode == ORC_MODE_PER) ? (L - 1 - L / 2) : (L - 2);
  for (int k = 0; k < Ho; ++k) {
    T* ra = a + (long)k * ap;
    T* rb = b + (long)k * bp;
    for (int j = 0; j < L; ++j) {
      long g = orc_ext_index(2L * k + j - pl, H, mode);
      const T c0 = f0[j], c1 = f1[j];
      if (g < 0) {
        if (j == 0) for (int c = 0; c < Wc; ++c) { ra[c] = c0 * (T)0; rb[c] = c1 * (T)0; }
        else for (int c = 0; c < Wc; ++c) { ra[c] = FMA(c0, (T)0, ra[c]); rb[c] = FMA(c1, (T)0, rb[c]); }
      } else {
        const T* src = in + g * inpitch;
        if (j == 0) for (int c = 0; c < Wc; ++c) { ra[c] = c0 * src[c]; rb[c] = c1 * src[c]; }
        else for (int c = 0; c < Wc; ++c) { ra[c] = FMA(c0, src[c], ra[c]); rb[c] = FMA(c1, src[c], rb[c]); }
      }
    }
  }
}

/* undecimated filter (colfilter/rowfilter) on a plane, odd or even L; acc: y += result */
static void FN(filt_plane_fast)(const T* x, int H, int W, long xpitch, T* y, long ypitch, const T* h, int L,
                                int symmetric, int along_w, int acc, T* e, T* t) {
  const int m = L / 2;
  const int mode = symmetric ? ORC_MODE_SYMMETRIC : ORC_MODE_ZERO;
  if (along_w) {
    const int Wo = W + 2 * m - L + 1;
    for (int r = 0; r < H; ++r) {
      FN(extend_line)(x + (long)r * xpitch, W, e, -m, Wo + L - 1, mode);
      T* yr = y + (long)r * ypitch;
      for (int n = 0; n < Wo; ++n) t[n] = h[0] * e[n];
      for (int j = 1; j < L; ++j) {
        const T c = h[j];
        for (int n = 0; n < Wo; ++n) t[n] = FMA(c, e[n + j], t[n]);
      }
      if (acc) for (int n = 0; n < Wo; ++n) yr[n] = yr[n] + t[n];
      else for (int n = 0; n < Wo; ++n) yr[n] = t[n];
    }
  } else {
    const int Ho = H + 2 * m - L + 1;
    for (int n = 0; n < Ho; ++n) {
      T* yr = y + (long)n * ypitch;
      for (int j = 0; j < L; ++j) {
        long g = orc_ext_index((long)n + j - m, H, mode);
        const T c = h[j];
        if (g < 0) {
          if (j == 0) for (int q = 0; q < W; ++q) t[q] = c * (T)0;
          else for (int q = 0; q < W; ++q) t[q] = FMA(c, (T)0, t[q]);
        } else {
          const T* src = x + g * xpitch;
          if (j == 0) for (int q = 0; q < W; ++q) t[q] = c * src[q];
          else for (int q = 0; q < W; ++q) t[q] = FMA(c, src[q], t[q]);
        }
      }
      if (acc) for (int q = 0; q < W; ++q) yr[q] = yr[q] + t[q];
      else for (int q = 0; q < W; ++q) yr[q] = t[q];
    }
  }
}

/* coldfilt/rowdfilt on a plane */
static void FN(dfilt_plane_fast)(const T* x, int H, int W, long xpitch, T* y, long ypitch, const T* ha,
                                 const T* hb, int m, int highpass, int along_w, T* e, T* ta, T* tb) {
  if (along_w) {
    const int Q = W / 4;
    for (int r = 0; r < H; ++r) {
      /* e[i] = x[sym(i + 2 - m)], i in [0, 4(Q-1) + 2(m-1) + 2) */
      FN(extend_line)(x + (long)r * xpitch, W, e, 2 - m, 4 * (Q - 1) + 2 * m, ORC_MODE_SYMMETRIC);
      T* yr = y + (long)r * ypitch;
      for (int q = 0; q < Q; ++q) { ta[q] = ha[0] * e[4 * q]; tb[q] = hb[0] * e[4 * q + 1]; }
      for (int j = 1; j < m; ++j) {
        const T ca = ha[j], cb = hb[j];
        for (int q = 0; q < Q; ++q) {
          ta[q] = FMA(ca, e[4 * q + 2 * j], ta[q]);
          tb[q] = FMA(cb, e[4 * q + 2 * j + 1], tb[q]);
        }
      }
      if (highpass) for (int q = 0; q < Q; ++q) { yr[2 * q] = tb[q]; yr[2 * q + 1] = ta[q]; }
      else for (int q = 0; q < Q; ++q) { yr[2 * q] = ta[q]; yr[2 * q + 1] = tb[q]; }
    }
  } else {
    const int Q = H / 4;
    for (int q = 0; q < Q; ++q) {
      T* ya = y + (long)(highpass ? 2 * q + 1 : 2 * q) * ypitch;
      T* yb = y + (long)(highpass ? 2 * q : 2 * q + 1) * ypitch;
      for (int j = 0; j < m; ++j) {
        const T* sa = x + orc_ext_index(4L * q + 2 * j + 2 - m, H, ORC_MODE_SYMMETRIC) * xpitch;
        const T* sb = x + orc_ext_index(4L * q + 2 * j + 3 - m, H, ORC_MODE_SYMMETRIC) * xpitch;
        const T ca = ha[j], cb = hb[j];
        if (j == 0) for (int c = 0; c < W; ++c) { ya[c] = ca * sa[c]; yb[c] = cb * sb[c]; }
        else for (int c = 0; c < W; ++c) { ya[c] = FMA(ca, sa[c], ya[c]); yb[c] = FMA(cb, sb[c], yb[c]); }
      }
    }
  }
}

/* colifilt/rowifilt on a plane; acc: y += result */
static void FN(ifilt_plane_fast)(const T* x, int H, int W, long xpitch, T* y, long ypitch, const T* ha,
                                 const T* hb, int m, int highpass, int along_w, int acc, T* e, T* t) {
  const int m2 = m / 2;
  int o[4], par[4];
  if (m2 % 2 == 0) {
    par[0] = 0; par[1] = 0; par[2] = 1; par[3] = 1;
    if (highpass) { o[0] = 1; o[1] = 0; o[2] = 3; o[3] = 2; } else { o[0] = 0; o[1] = 1; o[2] = 2; o[3] = 3; }
  } else {
    par[0] = 1; par[1] = 1; par[2] = 0; par[3] = 0;
    if (highpass) { o[0] = 2; o[1] = 1; o[2] = 2; o[3] = 1; } else { o[0] = 1; o[1] = 2; o[2] = 1; o[3] = 2; }
  }
  if (along_w) {
    const int Tn = W / 2;
    for (int r = 0; r < H; ++r) {
      /* e[i] = x[sym(i - m2)], i in [0, 2(Tn-1+m2-1) + 3 + 1) */
      FN(extend_line)(x + (long)r * xpitch, W, e, -m2, 2 * (Tn + m2) + 2, ORC_MODE_SYMMETRIC);
      T* yr = y + (long)r * ypitch;
      for (int s = 0; s < 4; ++s) {
        const T* h = (s & 1) ? hb : ha;
        for (int u = 0; u < Tn; ++u) t[u] = h[par[s]] * e[2 * u + o[s]];
        for (int j = 1; j < m2; ++j) {
          const T c = h[2 * j + par[s]];
          for (int u = 0; u < Tn; ++u) t[u] = FMA(c, e[2 * (u + j) + o[s]], t[u]);
        }
        if (acc) for (int u = 0; u < Tn; ++u) yr[4 * u + s] = yr[4 * u + s] + t[u];
        else for (int u = 0; u < Tn; ++u) yr[4 * u + s] = t[u];
      }
    }
  } else {
    const int Tn = H / 2;
    for (int u = 0; u < Tn; ++u)
      for (int s = 0; s < 4; ++s) {
        const T* h = (s & 1) ? hb : ha;
        T* yr = y + (4L * u + s) * ypitch;
        for (int j = 0; j < m2; ++j) {
          const T* src = x + orc_ext_index(2L * (u + j) + o[s] - m2, H, ORC_MODE_SYMMETRIC) * xpitch;
          const T c = h[2 * j + par[s]];
          if (j == 0) for (int q = 0; q < W; ++q) t[q] = c * src[q];
          else for (int q = 0; q < W; ++q) t[q] = FMA(c, src[q], t[q]);
        }
        if (acc) for (int q = 0; q < W; ++q) yr[q] = yr[q] + t[q];
        else for (int q = 0; q < W; ++q) yr[q] = t[q];
      }
  }
}

/* sfb along H, row-wise: lo/hi (K, Wc) (either may be null = zeros) -> y (Nout, Wc) */
static void FN(sfb_cols)(const T* lo, long lop, const T* hi, long hip, int K, int Wc, T* y, long yp, int Nout,
                         const T* g0, const T* g1, int L, int mode, T* ta, T* tb) {
  const int off = (mode == ORC_MODE_PER) ? (L / 2 - 1) : (L - 2);
  for (int n = 0; n < Nout; ++n) {
    long s = (long)n + off;
    long kmin = orc_floordiv(s - L + 2, 2), kmax = orc_floordiv(s, 2);
    int first = 1;
    for (long k = kmin; k <= kmax; ++k) {
      long kk = k;
      if (mode == ORC_MODE_PER) { kk = k % K; if (kk < 0) kk += K; }
      else if (k < 0 || k >= K) continue;
      const int t = (int)(s - 2 * k);
      const T c0 = g0[t], c1 = g1[t];
      const T* rl = lo ? lo + kk * lop : (const T*)0;
      const T* rh = hi ? hi + kk * hip : (const T*)0;
      if (first) {
        for (int c = 0; c < Wc; ++c) { ta[c] = (rl ? rl[c] : (T)0) * c0; tb[c] = (rh ? rh[c] : (T)0) * c1; }
        first = 0;
      } else {
        for (int c = 0; c < Wc; ++c) {
          ta[c] = FMA(rl ? rl[c] : (T)0, c0, ta[c]);
          tb[c] = FMA(rh ? rh[c] : (T)0, c1, tb[c]);
        }
      }
    }
    T* yr = y + (long)n * yp;
    if (first) for (int c = 0; c < Wc; ++c) yr[c] = (T)0 + (T)0;
    else for (int c = 0; c < Wc; ++c) yr[c] = ta[c] + tb[c];
  }
}

/* ---- plane-level passes (apply a line op along rows or columns of a (H,W) plane) ------------ */

static void FN(filt_plane)(const T* x, int H, int W, long xpitch, T* y, long ypitch, const T* h, int L,
                           int symmetric, int along_w, int acc) {
  if (orc_use_line_forms) {
    if (along_w) for (int r = 0; r < H; ++r) FN(filt_line)(x + r * xpitch, 1, W, y + r * ypitch, 1, h, L, symmetric, acc);
    else for (int c = 0; c < W; ++c) FN(filt_line)(x + c, xpitch, H, y + c, ypitch, h, L, symmetric, acc);
    return;
  }
  const size_t n = (size_t)(H > W ? H : W) + 2 * (size_t)L + 8;
  T* e = (T*)malloc(sizeof(T) * 2 * n);
  FN(filt_plane_fast)(x, H, W, xpitch, y, ypitch, h, L, symmetric, along_w, acc, e, e + n);
  free(e);
}
static void FN(dfilt_plane)(const T* x, int H, int W, long xpitch, T* y, long ypitch, const T* ha, const T* hb,
                            int m, int highpass, int along_w) {
  if (orc_use_line_forms) {
    if (along_w) for (int r = 0; r < H; ++r) FN(dfilt_line)(x + r * xpitch, 1, W, y + r * ypitch, 1, ha, hb, m, highpass);
    else for (int c = 0; c < W; ++c) FN(dfilt_line)(x + c, xpitch, H, y + c, ypitch, ha, hb, m, highpass);
    return;
  }
  const size_t n = (size_t)(H > W ? H : W) + 2 * (size_t)m + 8;
  T* e = (T*)malloc(sizeof(T) * 3 * n);
  FN(dfilt_plane_fast)(x, H, W, xpitch, y, ypitch, ha, hb, m, highpass, along_w, e, e + n, e + 2 * n);
  free(e);
}
static void FN(ifilt_plane)(const T* x, int H, int W, long xpitch, T* y, long ypitch, const T* ha, const T* hb,
                            int m, int highpass, int along_w, int acc) {
  if (orc_use_line_forms) {
    if (along_w) for (int r = 0; r < H; ++r) FN(ifilt_line)(x + r * xpitch, 1, W, y + r * ypitch, 1, ha, hb, m, highpass, acc);
    else for (int c = 0; c < W; ++c) FN(ifilt_line)(x + c, xpitch, H, y + c, ypitch, ha, hb, m, highpass, acc);
    return;
  }
  const size_t n = (size_t)(H > W ? H : W) + 2 * (size_t)m + 8;
  T* e = (T*)malloc(sizeof(T) * 2 * n);
  FN(ifilt_plane_fast)(x, H, W, xpitch, y, ypitch, ha, hb, m, highpass, along_w, acc, e, e + n);
  free(e);
}

/* ---- exported single-primitive entry points (unit tests vs the reference primitives) -------- */

/* colfilter (along_w=0) / rowfilter (along_w=1) on (planes,H,W) contiguous; y sized for Nout. */
int FN(orc_filter)(const T* x, T* y, int planes, int H, int W, const T* h, int L, int symmetric, int along_w) {
  const int m = L / 2, ext = 2 * m - L + 1;
  const int Ho = along_w ? H : H + ext, Wo = along_w ? W + ext : W;
#pragma omp parallel for schedule(static)
  for (int p = 0; p < planes; ++p)
    FN(filt_plane)(x + (long)p * H * W, H, W, W, y + (long)p * Ho * Wo, Wo, h, L, symmetric, along_w, 0);
  return 0;
}
/* coldfilt / rowdfilt */
int FN(orc_dfilt)(const T* x, T* y, int planes, int H, int W, const T* ha, const T* hb, int m, int highpass,
                  int along_w) {
  if ((along_w ? W : H) % 4 != 0) return ORC_ESIZE;
  const int Ho = along_w ? H : H / 2, Wo = along_w ? W / 2 : W;
#pragma omp parallel for schedule(static)
  for (int p = 0; p < planes; ++p)
    FN(dfilt_plane)(x + (long)p * H * W, H, W, W, y + (long)p * Ho * Wo, Wo, ha, hb, m, highpass, along_w);
  return 0;
}
/* colifilt / rowifilt */
int FN(orc_ifilt)(const T* x, T* y, int planes, int H, int W, const T* ha, const T* hb, int m, int highpass,
                  int along_w) {
  if ((along_w ? W : H) % 2 != 0) return ORC_ESIZE;
  const int Ho = along_w ? H : H * 2, Wo = along_w ? W * 2 : W;
#pragma omp parallel for schedule(static)
  for (int p = 0; p < planes; ++p)
    FN(ifilt_plane)(x + (long)p * H * W, H, W, W, y + (long)p * Ho * Wo, Wo, ha, hb, m, highpass, along_w, 0);
  return 0;
}

/* ---- 1-D DWT levels (AFB1D / SFB1D, dwt/lowlevel.py:368-424, 697-743): afb1d / sfb1d along the last dimension of
 * `rows` independent signals ----------------------------------------------------------------------- */
int FN(orc_dwt_afb1d)(const T* x, long long xpitch, int rows, int N, T* lo, T* hi, const T* f0, const T* f1, int L,
                      int mode) {
  if (!orc_mode_ok(mode)) return ORC_EMODE;
  if (N < 1 || L < 1 || rows < 0) return ORC_ESIZE;
  const int K = orc_coeff_len(N, L, mode);
#pragma omp parallel for schedule(static)
  for (int r = 0; r < rows; ++r)
    FN(afb_line)(x + (long long)r * xpitch, 1, N, lo + (long long)r * K, 1, hi + (long long)r * K, 1, K, f0, f1, L, mode);
  return 0;
}

int FN(orc_dwt_sfb1d)(const T* lo, const T* hi, int rows, int K, T* y, int Nout, const T* g0, const T* g1, int L,
                      int mode) {
  if (!orc_mode_ok(mode)) return ORC_EMODE;
  if (K < 1 || L < 1 || rows < 0 || Nout < 1 || Nout > orc_rec_len(K, L, mode)) return ORC_ESIZE;
#pragma omp parallel for schedule(static)
  for (int r = 0; r < rows; ++r)
    FN(sfb_line)(lo + (long long)r * K, 1, hi ? hi + (long long)r * K : (const T*)0, 1, K, y + (long long)r * Nout, 1,
                 Nout, g0, g1, L, mode);
  return 0;
}

/* ---- K1 / K2: DWT levels --------------------------------------------------------------------- */

/* AFB2D.forward, dwt/lowlevel.py:336-347: afb1d along W (dim=3) then along H (dim=2) on the
 * 2C-channel intermediate; channel order ll, lh, hl, hh with lh = low-W / high-H. */
int FN(orc_dwt_afb2d)(const T* x, long long xps, int xpitch, T* ll, long long llps, int llpitch, T* highs,
                      int planes, int H, int W, const T* fw_lo, const T* fw_hi, int Lw,
                      const T* fh_lo, const T* fh_hi, int Lh, int mode) {
  if (!orc_mode_ok(mode)) return ORC_EMODE;
  if (H < 1 || W < 1 || Lw < 1 || Lh < 1 || planes < 0) return ORC_ESIZE;
  const int Ho = orc_coeff_len(H, Lh, mode), Wo = orc_coeff_len(W, Lw, mode);
  int err = 0;
#pragma omp parallel for schedule(static)
  for (int p = 0; p < planes; ++p) {
    T* lo = (T*)malloc(sizeof(T) * (size_t)H * Wo * 2);
    if (!lo) { err = 1; continue; }
    T* hi = lo + (size_t)H * Wo;
    const T* xp = x + (long long)p * xps;
    T* llp = ll + (long long)p * llps;
    T* hp = highs + (long long)p * 3 * Ho * Wo;
    if (orc_use_line_forms) {
      for (int r = 0; r < H; ++r)
        FN(afb_line)(xp + (long)r * xpitch, 1, W, lo + (long)r * Wo, 1, hi + (long)r * Wo, 1, Wo, fw_lo, fw_hi, Lw, mode);
      for (int c = 0; c < Wo; ++c) {
        FN(afb_line)(lo + c, Wo, H, llp + c, llpitch, hp + c, Wo, Ho, fh_lo, fh_hi, Lh, mode);                     /* ll, lh */
        FN(afb_line)(hi + c, Wo, H, hp + (long)Ho * Wo + c, Wo, hp + 2L * Ho * Wo + c, Wo, Ho, fh_lo, fh_hi, Lh, mode); /* hl, hh */
      }
    } else {
      T* e = (T*)malloc(sizeof(T) * ((size_t)W + 2 * (size_t)Lw + 8));
      if (!e) { err = 1; free(lo); continue; }
      FN(afb_rows)(xp, xpitch, H, W, lo, hi, Wo, fw_lo, fw_hi, Lw, mode, e);
      FN(afb_cols)(lo, Wo, H, Wo, llp, llpitch, hp, Wo, Ho, fh_lo, fh_hi, Lh, mode);                       /* ll, lh */
      FN(afb_cols)(hi, Wo, H, Wo, hp + (long)Ho * Wo, Wo, hp + 2L * Ho * Wo, Wo, Ho, fh_lo, fh_hi, Lh, mode); /* hl, hh */
      free(e);
    }
    free(lo);
  }
  return err ? ORC_EINTERNAL : 0;
}

/* SFB2D.forward, dwt/lowlevel.py:671-680: sfb1d along H on (ll,lh) and (hl,hh), then along W. */
int FN(orc_dwt_sfb2d)(const T* ll, long long llps, int llpitch, const T* highs, T* y, long long yps, int ypitch,
                      int planes, int Hc, int Wc, int Ho, int Wo, const T* gh_lo, const T* gh_hi, int Lh,
                      const T* gw_lo, const T* gw_hi, int Lw, int mode) {
  if (!orc_mode_ok(mode)) return ORC_EMODE;
  const int Hn = orc_rec_len(Hc, Lh, mode), Wn = orc_rec_len(Wc, Lw, mode);
  if (Hc < 1 || Wc < 1 || Ho > Hn || Wo > Wn || Ho < 1 || Wo < 1) return ORC_ESIZE;
  int err = 0;
#pragma omp parallel for schedule(static)
  for (int p = 0; p < planes; ++p) {
    T* lo = (T*)malloc(sizeof(T) * (size_t)Hn * Wc * 2);
    if (!lo) { err = 1; continue; }
    T* hi = lo + (size_t)Hn * Wc;
    const T* llp = ll + (long long)p * llps;
    const T* hp = highs ? highs + (long long)p * 3 * Hc * Wc : (const T*)0;
    if (orc_use_line_forms) {
      for (int c = 0; c < Wc; ++c) {
        FN(sfb_line)(llp + c, llpitch, hp ? hp + c : 0, Wc, Hc, lo + c, Wc, Hn, gh_lo, gh_hi, Lh, mode);
        FN(sfb_line)(hp ? hp + (long)Hc * Wc + c : 0, Wc, hp ? hp + 2L * Hc * Wc + c : 0, Wc, Hc, hi + c, Wc, Hn,
                     gh_lo, gh_hi, Lh, mode);
      }
    } else {
      T* t = (T*)malloc(sizeof(T) * 2 * (size_t)Wc);
      if (!t) { err = 1; free(lo); continue; }
      FN(sfb_cols)(llp, llpitch, hp, Wc, Hc, Wc, lo, Wc, Hn, gh_lo, gh_hi, Lh, mode, t, t + Wc);
      FN(sfb_cols)(hp ? hp + (long)Hc * Wc : 0, Wc, hp ? hp + 2L * Hc * Wc : 0, Wc, Hc, Wc, hi, Wc, Hn, gh_lo, gh_hi,
                   Lh, mode, t, t + Wc);
      free(t);
    }
    T* yp = y + (long long)p * yps;
    for (int r = 0; r < Ho; ++r)
      FN(sfb_line)(lo + (long)r * Wc, 1, hi + (long)r * Wc, 1, Wc, yp + (long)r * ypitch, 1, Wo, gw_lo, gw_hi, Lw, mode);
    free(lo);
  }
  return err ? ORC_EINTERNAL : 0;
}

/* ---- DTCWT quad <-> complex packing ----------------------------------------------------------- */

/* transform_funcs.py:61-72 highs_to_orientations + dtcwt/lowlevel.py:243-260 q2c:
 * y/sqrt(2) first, then a-d, b+c (w1) and a+d, b-c (w2).  o1 = orientation slot of w1, o2 of w2. */
static void FN(q2c_store)(const T* y, int H, int W, long pitch, T* highs, const long long hs[6], long long base,
                          int o1, int o2) {
  const T r2 = (T)SQRT2_D;
  for (int i = 0; i < H / 2; ++i)
    for (int j = 0; j < W / 2; ++j) {
      T a = y[(2L * i) * pitch + 2 * j] / r2, b = y[(2L * i) * pitch + 2 * j + 1] / r2;
      T c = y[(2L * i + 1) * pitch + 2 * j] / r2, d = y[(2L * i + 1) * pitch + 2 * j + 1] / r2;
      long long q = base + i * hs[3] + j * hs[4];
      highs[q + o1 * hs[2]] = a - d;
      highs[q + o1 * hs[2] + hs[5]] = b + c;
      highs[q + o2 * hs[2]] = a + d;
      highs[q + o2 * hs[2] + hs[5]] = b - c;
    }
}

/* transform_funcs.py:75-95 orientations_to_highs + dtcwt/lowlevel.py:263-295 c2q. y is (2h, 2w). */
static void FN(c2q_load)(T* y, int h, int w, long pitch, const T* highs, const long long hs[6], long long base,
                         int o1, int o2) {
  const T r2 = (T)SQRT2_D;
  for (int i = 0; i < h; ++i)
    for (int j = 0; j < w; ++j) {
      long long q = base + i * hs[3] + j * hs[4];
      T w1r = highs[q + o1 * hs[2]], w1i = highs[q + o1 * hs[2] + hs[5]];
      T w2r = highs[q + o2 * hs[2]], w2i = highs[q + o2 * hs[2] + hs[5]];
      y[(2L * i) * pitch + 2 * j] = (w1r + w2r) / r2;
      y[(2L * i) * pitch + 2 * j + 1] = (w1i + w2i) / r2;
      y[(2L * i + 1) * pitch + 2 * j] = (w1i - w2i) / r2;
      y[(2L * i + 1) * pitch + 2 * j + 1] = (-w1r + w2r) / r2;
    }
}

/* ---- K3: FWD_J1.forward, transform_funcs.py:346-358 / fwd_j1 :98-121 -------------------------- */
int FN(orc_dtcwt_fwd_j1)(const T* x, long long xps, int xpitch, T* ll, long long llps, int llpitch, T* highs,
                         const long long hs[6], int N, int C, int H, int W, const T* h0, int L0, const T* h1, int L1,
                         int mode) {
  if (H < 2 || W < 2 || (H & 1) || (W & 1)) return ORC_ESIZE;
  if (!(L0 & 1) || !(L1 & 1)) return ORC_EFILTER;
  const int sym = (mode == ORC_MODE_SYMMETRIC);
  int err = 0;
#pragma omp parallel for schedule(static)
  for (int p = 0; p < N * C; ++p) {
    const size_t hw = (size_t)H * W;
    T* buf = (T*)malloc(sizeof(T) * hw * 3);
    if (!buf) { err = 1; continue; }
    T *lo = buf, *hi = buf + hw, *t = buf + 2 * hw;
    const T* xp = x + (long long)p * xps;
    T* llp = ll + (long long)p * llps;
    FN(filt_plane)(xp, H, W, xpitch, lo, W, h0, L0, sym, 1, 0);   /* lo = rowfilter(x, h0)  :107 */
    FN(filt_plane)(lo, H, W, W, llp, llpitch, h0, L0, sym, 0, 0); /* ll = colfilter(lo, h0) :109 */
    if (highs) {
      long long base = (long long)(p / C) * hs[0] + (long long)(p % C) * hs[1];
      FN(filt_plane)(xp, H, W, xpitch, hi, W, h1, L1, sym, 1, 0); /* hi = rowfilter(x, h1)  :108 */
      FN(filt_plane)(lo, H, W, W, t, W, h1, L1, sym, 0, 0);       /* lh = colfilter(lo, h1) :110 */
      FN(q2c_store)(t, H, W, W, highs, hs, base, 0, 5);           /* 15, 165 */
      FN(filt_plane)(hi, H, W, W, t, W, h0, L0, sym, 0, 0);       /* hl = colfilter(hi, h0) :112 */
      FN(q2c_store)(t, H, W, W, highs, hs, base, 2, 3);           /* 75, 105 */
      FN(filt_plane)(hi, H, W, W, t, W, h1, L1, sym, 0, 0);       /* hh = colfilter(hi, h1) :113 */
      FN(q2c_store)(t, H, W, W, highs, hs, base, 1, 4);           /* 45, 135 */
    }
    free(buf);
  }
  return err ? ORC_EINTERNAL : 0;
}

/* ---- K4: FWD_J2PLUS.forward, transform_funcs.py:380-392 / fwd_j2plus :226-249 ----------------- */
int FN(orc_dtcwt_fwd_j2plus)(const T* x, long long xps, int xpitch, T* ll, long long llps, int llpitch, T* highs,
                             const long long hs[6], int N, int C, int H, int W, const T* h0a, const T* h1a,
                             const T* h0b, const T* h1b, int m) {
  if (H < 4 || W < 4 || (H % 4) || (W % 4)) return ORC_ESIZE;
  if (m < 2 || (m & 1)) return ORC_EFILTER;
  int err = 0;
  const int H2 = H / 2, W2 = W / 2;
#pragma omp parallel for schedule(static)
  for (int p = 0; p < N * C; ++p) {
    const size_t hw2 = (size_t)H * W2;
    T* buf = (T*)malloc(sizeof(T) * (hw2 * 2 + (size_t)H2 * W2));
    if (!buf) { err = 1; continue; }
    T *lo = buf, *hi = buf + hw2, *t = buf + 2 * hw2;
    const T* xp = x + (long long)p * xps;
    T* llp = ll + (long long)p * llps;
    FN(dfilt_plane)(xp, H, W, xpitch, lo, W2, h0b, h0a, m, 0, 1);    /* lo = rowdfilt(x, h0b, h0a, False) :234 */
    FN(dfilt_plane)(lo, H, W2, W2, llp, llpitch, h0b, h0a, m, 0, 0); /* ll = coldfilt(lo, h0b, h0a, False) :237 */
    if (highs) {
      long long base = (long long)(p / C) * hs[0] + (long long)(p % C) * hs[1];
      FN(dfilt_plane)(xp, H, W, xpitch, hi, W2, h1b, h1a, m, 1, 1);  /* hi = rowdfilt(x, h1b, h1a, True)  :235 */
      FN(dfilt_plane)(lo, H, W2, W2, t, W2, h1b, h1a, m, 1, 0);      /* lh = coldfilt(lo, h1b, h1a, True) :238 */
      FN(q2c_store)(t, H2, W2, W2, highs, hs, base, 0, 5);
      FN(dfilt_plane)(hi, H, W2, W2, t, W2, h0b, h0a, m, 0, 0);      /* hl = coldfilt(hi, h0b, h0a, False) :239 */
      FN(q2c_store)(t, H2, W2, W2, highs, hs, base, 2, 3);
      FN(dfilt_plane)(hi, H, W2, W2, t, W2, h1b, h1a, m, 1, 0);      /* hh = coldfilt(hi, h1b, h1a, True) :240 */
      FN(q2c_store)(t, H2, W2, W2, highs, hs, base, 1, 4);
    }
    free(buf);
  }
  return err ? ORC_EINTERNAL : 0;
}

/* ---- K5: INV_J1.forward, transform_funcs.py:419-431 / inv_j1 :152-184 ------------------------- */
int FN(orc_dtcwt_inv_j1)(const T* ll, long long llps, int llpitch, const T* highs, const long long hs[6], T* y,
                         long long yps, int ypitch, int N, int C, int H, int W, const T* g0, int L0, const T* g1,
                         int L1, int mode) {
  if (H < 2 || W < 2 || (H & 1) || (W & 1)) return ORC_ESIZE;
  if (!(L0 & 1) || !(L1 & 1)) return ORC_EFILTER;
  if (!ll && !highs) return ORC_EARG;
  int err = 0;
#pragma omp parallel for schedule(static)
  for (int p = 0; p < N * C; ++p) {
    const size_t hw = (size_t)H * W;
    T* buf = (T*)malloc(sizeof(T) * hw * 3);
    if (!buf) { err = 1; continue; }
    T *lo = buf, *hi = buf + hw, *t = buf + 2 * hw;
    const T* llp = ll ? ll + (long long)p * llps : (const T*)0;
    T* yp = y + (long long)p * yps;
    if (!highs) {
      /* :159  y = rowfilter(colfilter(ll, g0), g0)  -- default mode='symmetric' whatever `mode` says */
      FN(filt_plane)(llp, H, W, llpitch, lo, W, g0, L0, 1, 0, 0);
      FN(filt_plane)(lo, H, W, W, yp, ypitch, g0, L0, 1, 1, 0);
    } else {
      const int sym = (mode == ORC_MODE_SYMMETRIC);
      long long base = (long long)(p / C) * hs[0] + (long long)(p % C) * hs[1];
      /* hi = colfilter(hh, g1) + colfilter(hl, g0)   :166/:178 */
      FN(c2q_load)(t, H / 2, W / 2, W, highs, hs, base, 1, 4); /* hh <- (45,135) :93 */
      FN(filt_plane)(t, H, W, W, hi, W, g1, L1, sym, 0, 0);
      FN(c2q_load)(t, H / 2, W / 2, W, highs, hs, base, 2, 3); /* hl <- (75,105) :92 */
      FN(filt_plane)(t, H, W, W, hi, W, g0, L0, sym, 0, 1);
      /* lo = colfilter(lh, g1) [+ colfilter(ll, g0)]   :167/:179 */
      FN(c2q_load)(t, H / 2, W / 2, W, highs, hs, base, 0, 5); /* lh <- (15,165) :91 */
      FN(filt_plane)(t, H, W, W, lo, W, g1, L1, sym, 0, 0);
      if (llp) FN(filt_plane)(llp, H, W, llpitch, lo, W, g0, L0, sym, 0, 1);
      /* y = rowfilter(hi, g1) + rowfilter(lo, g0)   :182 */
      FN(filt_plane)(hi, H, W, W, yp, ypitch, g1, L1, sym, 1, 0);
      FN(filt_plane)(lo, H, W, W, yp, ypitch, g0, L0, sym, 1, 1);
    }
    free(buf);
  }
  return err ? ORC_EINTERNAL : 0;
}

/* ---- K6: INV_J2PLUS.forward, transform_funcs.py:455-468 / inv_j2plus :279-307 ----------------- */
int FN(orc_dtcwt_inv_j2plus)(const T* ll, long long llps, int llpitch, const T* highs, const long long hs[6], T* y,
                             long long yps, int ypitch, int N, int C, int H, int W, const T* g0a, const T* g1a,
                             const T* g0b, const T* g1b, int m) {
  if (H < 2 || W < 2 || (H & 1) || (W & 1)) return ORC_ESIZE;
  if (m < 2 || (m & 1)) return ORC_EFILTER;
  if (!ll && !highs) return ORC_EARG;
  int err = 0;
#pragma omp parallel for schedule(static)
  for (int p = 0; p < N * C; ++p) {
    const size_t hw = (size_t)H * W;
    T* buf = (T*)malloc(sizeof(T) * hw * 5); /* lo, hi: (2H, W) each; t: (H, W) */
    if (!buf) { err = 1; continue; }
    T *lo = buf, *hi = buf + 2 * hw, *t = buf + 4 * hw;
    const T* llp = ll ? ll + (long long)p * llps : (const T*)0;
    T* yp = y + (long long)p * yps;
    if (!highs) {
      /* :286 y = rowifilt(colifilt(ll, g0b, g0a, False), g0b, g0a, False) */
      FN(ifilt_plane)(llp, H, W, llpitch, lo, W, g0b, g0a, m, 0, 0, 0);
      FN(ifilt_plane)(lo, 2 * H, W, W, yp, ypitch, g0b, g0a, m, 0, 1, 0);
    } else {
      long long base = (long long)(p / C) * hs[0] + (long long)(p % C) * hs[1];
      /* hi = colifilt(hh, g1b, g1a, True) + colifilt(hl, g0b, g0a, False)  :293/:299 */
      FN(c2q_load)(t, H / 2, W / 2, W, highs, hs, base, 1, 4);
      FN(ifilt_plane)(t, H, W, W, hi, W, g1b, g1a, m, 1, 0, 0);
      FN(c2q_load)(t, H / 2, W / 2, W, highs, hs, base, 2, 3);
      FN(ifilt_plane)(t, H, W, W, hi, W, g0b, g0a, m, 0, 0, 1);
      /* lo = colifilt(lh, g1b, g1a, True) [+ colifilt(ll, g0b, g0a, False)]  :295/:301 */
      FN(c2q_load)(t, H / 2, W / 2, W, highs, hs, base, 0, 5);
      FN(ifilt_plane)(t, H, W, W, lo, W, g1b, g1a, m, 1, 0, 0);
      if (llp) FN(ifilt_plane)(llp, H, W, llpitch, lo, W, g0b, g0a, m, 0, 0, 1);
      /* y = rowifilt(hi, g1b, g1a, True) + rowifilt(lo, g0b, g0a, False)  :305 */
      FN(ifilt_plane)(hi, 2 * H, W, W, yp, ypitch, g1b, g1a, m, 1, 1, 0);
      FN(ifilt_plane)(lo, 2 * H, W, W, yp, ypitch, g0b, g0a, m, 0, 1, 1);
    }
    free(buf);
  }
  return err ? ORC_EINTERNAL : 0;
}

/* ---- K7: ScatLayerj1_f.forward (combine_colour=False), scatternet/lowlevel.py:76-111 ----------- */
int FN(orc_scat_j1)(const T* x, T* z, T* dre, T* dim, int N, int C, int H, int W, const T* h0, int L0, const T* h1,
                    int L1, int mode, T magbias) {
  if (H < 2 || W < 2 || (H & 1) || (W & 1)) return ORC_ESIZE; /* :81 assert r%2 == c%2 == 0 */
  const int h = H / 2, w = W / 2;
  const size_t hw = (size_t)h * w;
  T* ll = (T*)malloc(sizeof(T) * (size_t)N * C * H * W);
  T* hg = (T*)malloc(sizeof(T) * (size_t)N * C * 12 * hw);
  if (!ll || !hg) { free(ll); free(hg); return ORC_EINTERNAL; }
  /* fwd_j1(x, h0o, h1o, False, 1, mode)  :87 -> reals/imags of shape (N, 6, C, h, w); keep ri outermost here */
  long long hs[6];
  hs[5] = (long long)N * 6 * C * hw; /* ri */
  hs[0] = 6LL * C * hw;              /* n  */
  hs[2] = (long long)C * hw;         /* o  */
  hs[1] = (long long)hw;             /* c  */
  hs[3] = w;
  hs[4] = 1;
  int rc = FN(orc_dtcwt_fwd_j1)(x, (long long)H * W, W, ll, (long long)H * W, W, hg, hs, N, C, H, W, h0, L0, h1, L1, mode);
  if (rc) { free(ll); free(hg); return rc; }
  const T b2 = (T)((double)magbias * (double)magbias); /* python float bias**2, cast on use */
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; ++n) {
    for (int c = 0; c < C; ++c) {
      const T* lp = ll + ((size_t)n * C + c) * H * W;
      T* z0 = z + (((size_t)n * 7 + 0) * C + c) * hw;
      for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) { /* F.avg_pool2d(ll, 2)  :88 */
          T s = lp[(2L * i) * W + 2 * j] + lp[(2L * i) * W + 2 * j + 1];
          s = s + lp[(2L * i + 1) * W + 2 * j];
          s = s + lp[(2L * i + 1) * W + 2 * j + 1];
          z0[(size_t)i * w + j] = s * (T)0.25;
        }
      for (int o = 0; o < 6; ++o) {
        const T* re = hg + (size_t)n * hs[0] + (size_t)o * hs[2] + (size_t)c * hs[1];
        const T* im = re + hs[5];
        T* zo = z + (((size_t)n * 7 + 1 + o) * C + c) * hw;
        size_t off6 = (((size_t)n * 6 + o) * C + c) * hw;
        for (size_t k = 0; k < hw; ++k) {
          T rr = re[k] * re[k];
          T ii = im[k] * im[k];
          T r = SQRT((rr + ii) + b2); /* :94 */
          if (dre) { dre[off6 + k] = re[k] / r; dim[off6 + k] = im[k] / r; } /* :96-99 */
          zo[k] = r - magbias; /* :104 */
        }
      }
    }
  }
  free(ll);
  free(hg);
  return 0;
}
