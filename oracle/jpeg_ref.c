/*
 * oracle/jpeg_ref.c — CPU restatement of the JPEG stripe encoder (CaptureSettings.output_mode = 0).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under selkies_b200/ may link, import or execute this file.
 *
 * What the reference fixes (its encoder lives in the out-of-tree pixelflux wheel): the WIRE format.  Every stripe the callback
 * delivers is  frame_id u16be | y_start u16be | a complete baseline JFIF file  (selkies.py:3116-3118 prepends 03 00; the client
 * reads frame id at offset 2, y_start at offset 4 and hands the rest to ImageDecoder as image/jpeg,
 * addons/selkies-web-core/selkies-ws-core.js:3166-3182, 2374-2393); quality comes from CaptureSettings.jpeg_quality /
 * paint_over_jpeg_quality (selkies.py:3209-3212).
 *
 * The algorithm restated is ITU-T T.81 baseline sequential DCT with the IJG conventions every JFIF producer follows:
 *   - forward DCT: the accurate integer LL&M transform (IJG "islow", 13-bit constants, results scaled by 8)
 *   - quantisation: Annex K tables scaled by the IJG quality rule; round-half-away division
 *   - Huffman: the Annex K.3-K.6 tables; DC differences, (run,size) AC symbols, ZRL, EOB; FF byte stuffing; 1-padding
 *   - 4:2:0 interleaved scan (MCU = 4 Y blocks, Cb, Cr), or a single-component scan
 * Pin (tests/test_jpeg_oracle.py): in single-component mode the output is compared BYTE FOR BYTE with libjpeg-turbo (through
 * cv2.imencode on the same grey image and quality) — an independent implementation of the same conventions; 4:2:0 streams are
 * decoded by libjpeg-turbo and must come back within the PSNR the quality implies.  Colour: JFIF full-range BT.601
 * (oracle/csc_ref.c matrix 1), 2x2 box chroma.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const uint8_t std_luma_q[64] = {
  16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
  18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99 };
static const uint8_t std_chroma_q[64] = {
  17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
  99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99 };
static const uint8_t zigzag[64] = {   /* scan position -> natural (row-major) index */
  0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
  35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };
/* Annex K.3-K.6: number of codes of each length 1..16, then the symbols */
static const uint8_t dc_luma_bits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t dc_chroma_bits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t dc_vals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t ac_luma_bits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t ac_luma_vals[162] = {
  0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08,
  0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
  0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
  0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
  0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
  0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
  0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa };
static const uint8_t ac_chroma_bits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t ac_chroma_vals[162] = {
  0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91,
  0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
  0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
  0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
  0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
  0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
  0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa };

typedef struct { uint16_t code[256]; uint8_t size[256]; } huff_t;
static void build_huff(huff_t* h, const uint8_t bits[16], const uint8_t* vals) {   /* Annex C */
  memset(h, 0, sizeof *h);
  int code = 0, k = 0;
  for (int len = 1; len <= 16; len++) {
    for (int i = 0; i < bits[len - 1]; i++, k++) { h->code[vals[k]] = (uint16_t)code++; h->size[vals[k]] = (uint8_t)len; }
    code <<= 1;
  }
}

/* quantiser table for a quality 1..100 (IJG rule), natural order */
void b2v_ref_jpeg_qtable(int quality, int chroma, uint8_t out[64]) {
  if (quality < 1) quality = 1;
  if (quality > 100) quality = 100;
  const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;
  const uint8_t* base = chroma ? std_chroma_q : std_luma_q;
  for (int i = 0; i < 64; i++) {
    int v = (base[i] * scale + 50) / 100;
    out[i] = (uint8_t)(v < 1 ? 1 : v > 255 ? 255 : v);
  }
}

#define CB 13
#define P1 2
#define DS(x, n) (((x) + (1 << ((n) - 1))) >> (n))
/* accurate integer forward DCT (LL&M), in place on 64 level-shifted samples, natural order; output scaled by 8 */
void b2v_ref_jpeg_fdct(int d[64]) {
  for (int pass = 0; pass < 2; pass++) {
    for (int i = 0; i < 8; i++) {
      int* p = pass == 0 ? d + 8 * i : d + i;
      const int st = pass == 0 ? 1 : 8;
      int t0 = p[0] + p[7 * st], t7 = p[0] - p[7 * st], t1 = p[st] + p[6 * st], t6 = p[st] - p[6 * st];
      int t2 = p[2 * st] + p[5 * st], t5 = p[2 * st] - p[5 * st], t3 = p[3 * st] + p[4 * st], t4 = p[3 * st] - p[4 * st];
      int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
      if (pass == 0) { p[0] = (t10 + t11) << P1; p[4 * st] = (t10 - t11) << P1; }
      else { p[0] = DS(t10 + t11, P1); p[4 * st] = DS(t10 - t11, P1); }
      const int sh = pass == 0 ? CB - P1 : CB + P1;
      int z1 = (t12 + t13) * 4433;
      p[2 * st] = DS(z1 + t13 * 6270, sh);
      p[6 * st] = DS(z1 - t12 * 15137, sh);
      z1 = t4 + t7; int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7, z5 = (z3 + z4) * 9633;
      t4 *= 2446; t5 *= 16819; t6 *= 25172; t7 *= 12299;
      z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
      z3 += z5; z4 += z5;
      p[7 * st] = DS(t4 + z1 + z3, sh); p[5 * st] = DS(t5 + z2 + z4, sh);
      p[3 * st] = DS(t6 + z2 + z3, sh); p[st] = DS(t7 + z1 + z4, sh);
    }
  }
}

/* 8x8 block at (bx,by) of a plane (pixel step `step`: 1 for Y, 2 for interleaved CbCr) -> quantised levels in SCAN order */
static void block_levels(const uint8_t* plane, int pitch, int step, int bx, int by, const uint8_t q[64], int16_t lv[64]) {
  int d[64];
  for (int y = 0; y < 8; y++)
    for (int x = 0; x < 8; x++) d[8 * y + x] = (int)plane[(size_t)(by * 8 + y) * pitch + (bx * 8 + x) * step] - 128;
  b2v_ref_jpeg_fdct(d);
  for (int k = 0; k < 64; k++) {
    const int n = zigzag[k], qv = (int)q[n] << 3;
    int t = d[n];
    if (t < 0) { t = -t; t += qv >> 1; t /= qv; t = -t; } else { t += qv >> 1; t /= qv; }
    lv[k] = (int16_t)t;
  }
}

typedef struct { uint8_t* out; size_t pos; uint32_t acc; int n; } jbw_t;
static void jb_put(jbw_t* b, int len, uint32_t v) {
  for (int i = len - 1; i >= 0; i--) {
    b->acc = (b->acc << 1) | ((v >> i) & 1u);
    if (++b->n == 8) { b->out[b->pos++] = (uint8_t)b->acc; if ((b->acc & 255u) == 255u) b->out[b->pos++] = 0; b->acc = 0; b->n = 0; }
  }
}
static int bitlen(int v) { int n = 0; while (v) { n++; v >>= 1; } return n; }
static void code_block(jbw_t* b, const int16_t lv[64], int* last_dc, const huff_t* dc, const huff_t* ac) {
  int t = lv[0] - *last_dc, t2 = t;
  *last_dc = lv[0];
  if (t < 0) { t = -t; t2--; }
  int nb = bitlen(t);
  jb_put(b, dc->size[nb], dc->code[nb]);
  if (nb) jb_put(b, nb, (uint32_t)t2 & ((1u << nb) - 1));
  int r = 0;
  for (int k = 1; k < 64; k++) {
    t = lv[k];
    if (t == 0) { r++; continue; }
    while (r > 15) { jb_put(b, ac->size[0xF0], ac->code[0xF0]); r -= 16; }
    t2 = t;
    if (t < 0) { t = -t; t2--; }
    nb = bitlen(t);
    jb_put(b, ac->size[(r << 4) + nb], ac->code[(r << 4) + nb]);
    jb_put(b, nb, (uint32_t)t2 & ((1u << nb) - 1));
    r = 0;
  }
  if (r > 0) jb_put(b, ac->size[0], ac->code[0]);
}

static size_t put_marker(uint8_t* o, int m, int len) { o[0] = 0xFF; o[1] = (uint8_t)m; o[2] = (uint8_t)(len >> 8); o[3] = (uint8_t)len; return 4; }
static size_t put_dht(uint8_t* o, int tc_th, const uint8_t bits[16], const uint8_t* vals, int nvals) {
  size_t n = put_marker(o, 0xC4, 2 + 1 + 16 + nvals);
  o[n++] = (uint8_t)tc_th; memcpy(o + n, bits, 16); n += 16; memcpy(o + n, vals, nvals); n += nvals;
  return n;
}

/* everything in front of the entropy-coded segment; ncomp = 1 (grey) or 3 (4:2:0) */
size_t b2v_ref_jpeg_header(uint8_t* o, int w, int h, int ncomp, int quality) {
  size_t n = 0;
  uint8_t q[64];
  o[n++] = 0xFF; o[n++] = 0xD8;
  n += put_marker(o + n, 0xE0, 16); memcpy(o + n, "JFIF\0\1\1\0\0\1\0\1\0\0", 14); n += 14;
  for (int t = 0; t < (ncomp == 3 ? 2 : 1); t++) {
    b2v_ref_jpeg_qtable(quality, t, q);
    n += put_marker(o + n, 0xDB, 67); o[n++] = (uint8_t)t;
    for (int k = 0; k < 64; k++) o[n++] = q[zigzag[k]];
  }
  n += put_marker(o + n, 0xC0, 8 + 3 * ncomp);
  o[n++] = 8; o[n++] = (uint8_t)(h >> 8); o[n++] = (uint8_t)h; o[n++] = (uint8_t)(w >> 8); o[n++] = (uint8_t)w; o[n++] = (uint8_t)ncomp;
  for (int c = 0; c < ncomp; c++) { o[n++] = (uint8_t)(c + 1); o[n++] = (uint8_t)(ncomp == 3 && c == 0 ? 0x22 : 0x11); o[n++] = (uint8_t)(c ? 1 : 0); }
  n += put_dht(o + n, 0x00, dc_luma_bits, dc_vals, 12);
  n += put_dht(o + n, 0x10, ac_luma_bits, ac_luma_vals, 162);
  if (ncomp == 3) { n += put_dht(o + n, 0x01, dc_chroma_bits, dc_vals, 12); n += put_dht(o + n, 0x11, ac_chroma_bits, ac_chroma_vals, 162); }
  n += put_marker(o + n, 0xDA, 6 + 2 * ncomp);
  o[n++] = (uint8_t)ncomp;
  for (int c = 0; c < ncomp; c++) { o[n++] = (uint8_t)(c + 1); o[n++] = (uint8_t)(c ? 0x11 : 0x00); }
  o[n++] = 0; o[n++] = 63; o[n++] = 0;
  return n;
}

/*
 * One JFIF file from planar data: y = luma rows (pitch bytes), uv = interleaved Cb,Cr rows at half resolution (same pitch) or
 * NULL for a single-component image.  w,h: the size written to SOF0; the planes must be readable up to the next multiple of
 * 16 (8 when grey) in both directions (the caller pads by replication, as csc_ref.c does).  Returns the file size.
 */
int64_t b2v_ref_jpeg_encode(const uint8_t* y, const uint8_t* uv, int pitch, int w, int h, int quality, uint8_t* out) {
  huff_t hdc[2], hac[2];
  build_huff(&hdc[0], dc_luma_bits, dc_vals); build_huff(&hac[0], ac_luma_bits, ac_luma_vals);
  build_huff(&hdc[1], dc_chroma_bits, dc_vals); build_huff(&hac[1], ac_chroma_bits, ac_chroma_vals);
  uint8_t ql[64], qc[64];
  b2v_ref_jpeg_qtable(quality, 0, ql); b2v_ref_jpeg_qtable(quality, 1, qc);
  const int ncomp = uv ? 3 : 1;
  size_t n = b2v_ref_jpeg_header(out, w, h, ncomp, quality);
  jbw_t b = { out + n, 0, 0, 0 };
  int16_t lv[64];
  int dc[3] = {0, 0, 0};
  if (ncomp == 1) {
    for (int by = 0; by < (h + 7) / 8; by++)
      for (int bx = 0; bx < (w + 7) / 8; bx++) { block_levels(y, pitch, 1, bx, by, ql, lv); code_block(&b, lv, &dc[0], &hdc[0], &hac[0]); }
  } else {
    for (int my = 0; my < (h + 15) / 16; my++)
      for (int mx = 0; mx < (w + 15) / 16; mx++) {
        for (int k = 0; k < 4; k++) { block_levels(y, pitch, 1, 2 * mx + (k & 1), 2 * my + (k >> 1), ql, lv); code_block(&b, lv, &dc[0], &hdc[0], &hac[0]); }
        block_levels(uv, pitch, 2, mx, my, qc, lv); code_block(&b, lv, &dc[1], &hdc[1], &hac[1]);
        block_levels(uv + 1, pitch, 2, mx, my, qc, lv); code_block(&b, lv, &dc[2], &hdc[1], &hac[1]);
      }
  }
  if (b.n) jb_put(&b, 8 - b.n, (1u << (8 - b.n)) - 1);      /* pad the last byte with 1-bits */
  n += b.pos;
  out[n++] = 0xFF; out[n++] = 0xD9;
  return (int64_t)n;
}
