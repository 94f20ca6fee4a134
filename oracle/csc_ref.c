/*
 * oracle/csc_ref.c — CPU restatement of the colour-convert(+scale) stage.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under selkies_b200/ may link, import or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker.
 *
 * PARITY UNPINNED: the reference (selkies @1a9cd02b) holds no implementation of
 * this arithmetic.  Its conversion lives in the out-of-tree `pixelflux` wheel
 * (pyproject.toml:37; call site media_pipeline.py:299-300) or, in the legacy
 * design BASELINE.json names, in GStreamer 1.24.12 `videoconvert`
 * (addons/gstreamer/Dockerfile:85,93; docs/component.md:338-344 "BGRx to I420
 * or NV12").  Neither source is under /root/reference and the reference has no
 * tests or golden vectors.  This file therefore restates
 * the published algorithm — ITU-R BT.709 limited-range 8-bit Y'CbCr — as an
 * exact integer spec, pinned by the BT.709 known-answer colours
 * (tests/test_csc_oracle.py) and by <= 1 LSB agreement with a float64
 * evaluation of the BT.709 matrix.
 *
 * Spec:
 *   Y  = 16  + ((2991 R + 10064 G + 1016 B + 2^13) >> 14)          per pixel
 *   Cb = 128 + ((-1649 ΣR - 5547 ΣG + 7196 ΣB + 2^15) >> 16)       per 2x2 block
 *   Cr = 128 + (( 7196 ΣR - 6536 ΣG -  660 ΣB + 2^15) >> 16)       (Σ over the 4 pixels)
 *   >> is an arithmetic (floor) shift.  Coefficients are rint(k * 2^14) of the
 *   BT.709 limited-range matrix; chroma is the matrix applied to the 2x2 box
 *   sum with a single rounding (centre-sited 4:2:0).
 *   Output NV12: Y plane (pitch = out width) then interleaved Cb,Cr rows.
 *
 * Scaling (when dst != src), applied to B,G,R before the matrix:
 *   pos(d, s_n, d_n) = clamp(floor(((2d+1) * s_n * 2^15) / d_n) - 2^15, 0, (s_n-1) << 16)
 *   i0 = pos >> 16, i1 = min(i0+1, s_n-1), f = (pos >> 8) & 255
 *   v = ((p00*(256-fx) + p01*fx) * (256-fy) + (p10*(256-fx) + p11*fx) * fy + 2^15) >> 16
 *
 * Padding: when coded size (multiple of 16) exceeds dst size, output pixel
 * (x,y) of the coded frame is the converted pixel (min(x,dst_w-1), min(y,dst_h-1))
 * — i.e. edge replication in the scaled-BGR domain (chroma sums then see the
 * replicated pixels).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* host threads used by the OpenMP loops of the oracle (csc_ref.c, h264_ref.c); 0 = all */
void b2v_ref_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
#else
  (void)n;
#endif
}
int b2v_ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

#define KYR 2991
#define KYG 10064
#define KYB 1016
#define KUR (-1649)
#define KUG (-5547)
#define KUB 7196
#define KVR 7196
#define KVG (-6536)
#define KVB (-660)

static inline int asr(int v, int s) { /* floor shift, independent of compiler >> semantics */
  return v >= 0 ? (v >> s) : -((-v + (1 << s) - 1) >> s);
}

/* source position for destination index d, 16.16 fixed point */
static int64_t scale_pos(int d, int sn, int dn) {
  int64_t p = (((int64_t)(2 * d + 1) * sn) << 15) / dn - (1 << 15);
  int64_t hi = (int64_t)(sn - 1) << 16;
  if (p < 0) p = 0;
  if (p > hi) p = hi;
  return p;
}

typedef struct { int i0, i1, f; } tap_t;

static void make_taps(tap_t* t, int dn, int sn) {
  for (int d = 0; d < dn; d++) {
    if (dn == sn) { t[d].i0 = d; t[d].i1 = d; t[d].f = 0; continue; }
    int64_t p = scale_pos(d, sn, dn);
    t[d].i0 = (int)(p >> 16);
    t[d].i1 = t[d].i0 + 1 < sn ? t[d].i0 + 1 : sn - 1;
    t[d].f = (int)((p >> 8) & 255);
  }
}

/* scaled B,G,R of destination pixel (x,y) (already clamped to dst range) */
static inline void fetch_bgr(const uint8_t* src, int stride, const tap_t* tx, const tap_t* ty,
                             int x, int y, int bgr[3]) {
  const uint8_t* r0 = src + (size_t)ty[y].i0 * stride;
  const uint8_t* r1 = src + (size_t)ty[y].i1 * stride;
  int x0 = tx[x].i0 * 4, x1 = tx[x].i1 * 4, fx = tx[x].f, fy = ty[y].f;
  if (fx == 0 && fy == 0) {   /* exact tap: the interpolation formula reduces to p00 */
    bgr[0] = r0[x0]; bgr[1] = r0[x0 + 1]; bgr[2] = r0[x0 + 2];
    return;
  }
  for (int c = 0; c < 3; c++) {
    int top = r0[x0 + c] * (256 - fx) + r0[x1 + c] * fx;
    int bot = r1[x0 + c] * (256 - fx) + r1[x1 + c] * fx;
    bgr[c] = (top * (256 - fy) + bot * fy + (1 << 15)) >> 16;
  }
}

/*
 * bgra:   src_h rows of src_stride bytes, B,G,R,A byte order (A ignored)
 * out_y:  coded_h rows of coded_w bytes;  out_uv: coded_h/2 rows of coded_w bytes (Cb,Cr pairs)
 * dst_w/dst_h: visible (scaled) size, even; coded_w/coded_h >= dst, even.
 */
/* matrix 0: BT.709 limited range (the H.264 path; the spec in the header).  matrix 1: JFIF full-range BT.601 (the JPEG stripe
 * path, oracle/jpeg_ref.c): Y = (4899 R + 9617 G + 1868 B + 2^13) >> 14, Cb = 128 + ((-2765 SR - 5427 SG + 8191 SB + 2^15) >> 16),
 * Cr = 128 + ((8191 SR - 6860 SG - 1332 SB + 2^15) >> 16); same scaling, padding and 2x2 box sum.  (0.5 is coded as 8191, not
 * 8192: a saturated red or blue would otherwise reach 256; neutral greys still give exactly 128.) */
static const int MATRIX[2][10] = {
  { KYR, KYG, KYB, KUR, KUG, KUB, KVR, KVG, KVB, 16 },
  { 4899, 9617, 1868, -2765, -5427, 8191, 8191, -6860, -1332, 0 } };

int b2v_ref_csc_nv12_m(const uint8_t* bgra, int src_w, int src_h, int src_stride,
                       int dst_w, int dst_h, int coded_w, int coded_h,
                       uint8_t* out_y, uint8_t* out_uv, int matrix);
int b2v_ref_csc_nv12(const uint8_t* bgra, int src_w, int src_h, int src_stride,
                     int dst_w, int dst_h, int coded_w, int coded_h,
                     uint8_t* out_y, uint8_t* out_uv) {
  return b2v_ref_csc_nv12_m(bgra, src_w, src_h, src_stride, dst_w, dst_h, coded_w, coded_h, out_y, out_uv, 0);
}
int b2v_ref_csc_nv12_m(const uint8_t* bgra, int src_w, int src_h, int src_stride,
                       int dst_w, int dst_h, int coded_w, int coded_h,
                       uint8_t* out_y, uint8_t* out_uv, int matrix) {
  if (matrix < 0 || matrix > 1) return -1;
  const int* M = MATRIX[matrix];
  if (src_w < 2 || src_h < 2 || (dst_w & 1) || (dst_h & 1) || (coded_w & 1) || (coded_h & 1) ||
      coded_w < dst_w || coded_h < dst_h)
    return -1;
  tap_t* tx = (tap_t*)malloc(sizeof(tap_t) * dst_w);
  tap_t* ty = (tap_t*)malloc(sizeof(tap_t) * dst_h);
  if (!tx || !ty) { free(tx); free(ty); return -3; }
  make_taps(tx, dst_w, src_w);
  make_taps(ty, dst_h, src_h);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < coded_h; y += 2) {
    for (int x = 0; x < coded_w; x += 2) {
      int sb = 0, sg = 0, sr = 0;
      for (int dy = 0; dy < 2; dy++)
        for (int dx = 0; dx < 2; dx++) {
          int px = x + dx < dst_w ? x + dx : dst_w - 1;
          int py = y + dy < dst_h ? y + dy : dst_h - 1;
          int bgr[3];
          fetch_bgr(bgra, src_stride, tx, ty, px, py, bgr);
          out_y[(size_t)(y + dy) * coded_w + x + dx] =
              (uint8_t)(M[9] + ((M[0] * bgr[2] + M[1] * bgr[1] + M[2] * bgr[0] + (1 << 13)) >> 14));
          sb += bgr[0]; sg += bgr[1]; sr += bgr[2];
        }
      out_uv[(size_t)(y >> 1) * coded_w + x]     = (uint8_t)(128 + asr(M[3] * sr + M[4] * sg + M[5] * sb + (1 << 15), 16));
      out_uv[(size_t)(y >> 1) * coded_w + x + 1] = (uint8_t)(128 + asr(M[6] * sr + M[7] * sg + M[8] * sb + (1 << 15), 16));
    }
  }
  free(tx); free(ty);
  return 0;
}
