// h264_intra.cu — IDR pictures: Intra16x16 macroblocks (ITU-T H.264 8.3.3 luma, 8.3.4 chroma).
//
// One warp per macroblock ROW: intra prediction needs the reconstructed left neighbour, so the
// macroblocks of a row are a serial chain; rows of the same slice additionally wait for the row above
// (wavefront, one macroblock of lag) through a progress counter in global memory.  With the default
// slice_rows = 1 every row is its own slice and all rows run fully in parallel.
// Encoder decisions: DESIGN.md §5.2; CPU restatement: oracle/h264_ref.c encode_intra_mb().
#include "h264_common.cuh"
#include "h264_kernels.h"

namespace b2v {

__device__ __forceinline__ uint32_t rep4(int v) { return (uint32_t)v * 0x01010101u; }

struct IntraNb {
  uint8_t top_y[16]; uint8_t left_y[16];
  uint8_t top_uv[16]; uint8_t left_uv[16];   // interleaved Cb,Cr: [x*2+c] / [y*2+c]
  int tl_y, tl_u, tl_v;
  int cdc[8];                                // chroma DC prediction per (comp*4 + blk)
};

__global__ void __launch_bounds__(32) k_intra_rows(FrameCtx f) {
  __shared__ __align__(16) MbTile t;
  __shared__ __align__(16) IntraNb nb;
  const int lane = threadIdx.x, mby = blockIdx.x;
  const int qp = frame_qp(f);
  const bool has_top = top_in_slice(f, mby);
  const size_t ysz = (size_t)f.cw * f.ch;
  const uint8_t* cur_y = f.cur; const uint8_t* cur_uv = f.cur + ysz;
  uint8_t* rec_y = f.recon; uint8_t* rec_uv = f.recon + ysz;
  const int r8 = lane >> 1, c8 = (lane & 1) * 8;            // this lane's 8 luma pixels
  const int rc4 = lane >> 2, cc4 = (lane & 3) * 4;          // this lane's 4 interleaved chroma bytes

  for (int mbx = 0; mbx < f.mbw; mbx++) {
    const bool has_left = mbx > 0;
    if (has_top) {       // wavefront: the row above must have finished macroblock mbx
      if (lane == 0) { while (*((volatile int*)&f.progress[mby - 1]) < mbx + 1) { } __threadfence(); }
      __syncwarp();
    }
    // ---- load current macroblock and neighbours --------------------------------------------------
    {
      const uint2 v = *reinterpret_cast<const uint2*>(cur_y + (size_t)(mby * 16 + r8) * f.cw + mbx * 16 + c8);
      *reinterpret_cast<uint2*>(&t.cur_y[r8][c8]) = v;
      if (lane < 16) {
        const uint2 w = *reinterpret_cast<const uint2*>(cur_uv + (size_t)(mby * 8 + r8) * f.cw + mbx * 16 + c8);
        *reinterpret_cast<uint2*>(&t.cur_uv[r8][c8]) = w;
      }
      if (has_top) {
        if (lane < 16) nb.top_y[lane] = __ldcg(rec_y + (size_t)(mby * 16 - 1) * f.cw + mbx * 16 + lane);
        else nb.top_uv[lane - 16] = __ldcg(rec_uv + (size_t)(mby * 8 - 1) * f.cw + mbx * 16 + (lane - 16));
        if (has_left && lane == 0) {
          nb.tl_y = __ldcg(rec_y + (size_t)(mby * 16 - 1) * f.cw + mbx * 16 - 1);
          nb.tl_u = __ldcg(rec_uv + (size_t)(mby * 8 - 1) * f.cw + mbx * 16 - 2);
          nb.tl_v = __ldcg(rec_uv + (size_t)(mby * 8 - 1) * f.cw + mbx * 16 - 1);
        }
      }
    }
    __syncwarp();
    // chroma DC predictors (8.3.4.1-3), one (component, block) per lane 0..7
    if (lane < 8) {
      const int c = lane >> 2, b = lane & 3, bx = (b & 1) * 4, by = (b >> 1) * 4;
      int st = 0, sl = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) { st += nb.top_uv[(bx + i) * 2 + c]; sl += nb.left_uv[(by + i) * 2 + c]; }
      int dc;
      if (b == 0 || b == 3) dc = has_top && has_left ? (st + sl + 4) >> 3 : has_top ? (st + 2) >> 2 : has_left ? (sl + 2) >> 2 : 128;
      else if (b == 1) dc = has_top ? (st + 2) >> 2 : has_left ? (sl + 2) >> 2 : 128;
      else dc = has_left ? (sl + 2) >> 2 : has_top ? (st + 2) >> 2 : 128;
      nb.cdc[lane] = dc;
    }
    // ---- luma mode decision: key = SAD*4 + mode over available modes -------------------------------
    const uint32_t cy0 = *reinterpret_cast<const uint32_t*>(&t.cur_y[r8][c8]);
    const uint32_t cy1 = *reinterpret_cast<const uint32_t*>(&t.cur_y[r8][c8 + 4]);
    int sum_t = 0, sum_l = 0, H = 0, V = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) { sum_t += nb.top_y[i]; sum_l += nb.left_y[i]; }
#pragma unroll
    for (int i = 0; i < 8; i++) {
      H += (i + 1) * ((int)nb.top_y[8 + i] - (i == 7 ? nb.tl_y : (int)nb.top_y[6 - i]));
      V += (i + 1) * ((int)nb.left_y[8 + i] - (i == 7 ? nb.tl_y : (int)nb.left_y[6 - i]));
    }
    const int dc_y = has_top && has_left ? (sum_t + sum_l + 16) >> 5 : has_top ? (sum_t + 8) >> 4 : has_left ? (sum_l + 8) >> 4 : 128;
    const int pa = 16 * ((int)nb.left_y[15] + (int)nb.top_y[15]), pb = (5 * H + 32) >> 6, pc = (5 * V + 32) >> 6;
    uint32_t pv0 = *reinterpret_cast<const uint32_t*>(&nb.top_y[c8]), pv1 = *reinterpret_cast<const uint32_t*>(&nb.top_y[c8 + 4]);
    uint32_t ph = rep4(nb.left_y[r8]), pd = rep4(dc_y), pp0 = 0, pp1 = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      pp0 |= (uint32_t)clip255((pa + pb * (c8 + j - 7) + pc * (r8 - 7) + 16) >> 5) << (8 * j);
      pp1 |= (uint32_t)clip255((pa + pb * (c8 + 4 + j - 7) + pc * (r8 - 7) + 16) >> 5) << (8 * j);
    }
    int best_key = 0x7fffffff, best_mode = 2;
    {
      int s;
      if (has_top) { s = __reduce_add_sync(FULL, __vsadu4(cy0, pv0) + __vsadu4(cy1, pv1)); if (s * 4 + 0 < best_key) { best_key = s * 4 + 0; best_mode = 0; } }
      if (has_left) { s = __reduce_add_sync(FULL, __vsadu4(cy0, ph) + __vsadu4(cy1, ph)); if (s * 4 + 1 < best_key) { best_key = s * 4 + 1; best_mode = 1; } }
      s = __reduce_add_sync(FULL, __vsadu4(cy0, pd) + __vsadu4(cy1, pd)); if (s * 4 + 2 < best_key) { best_key = s * 4 + 2; best_mode = 2; }
      if (has_top && has_left) { s = __reduce_add_sync(FULL, __vsadu4(cy0, pp0) + __vsadu4(cy1, pp1)); if (s * 4 + 3 < best_key) { best_key = s * 4 + 3; best_mode = 3; } }
    }
    {
      uint32_t p0 = best_mode == 0 ? pv0 : best_mode == 1 ? ph : best_mode == 2 ? pd : pp0;
      uint32_t p1 = best_mode == 0 ? pv1 : best_mode == 1 ? ph : best_mode == 2 ? pd : pp1;
      *reinterpret_cast<uint32_t*>(&t.pred_y[r8][c8]) = p0;
      *reinterpret_cast<uint32_t*>(&t.pred_y[r8][c8 + 4]) = p1;
    }
    __syncwarp();   // nb.cdc visible
    // ---- chroma mode decision: key = (SAD Cb + SAD Cr)*4 + mode ------------------------------------
    const uint32_t cc = *reinterpret_cast<const uint32_t*>(&t.cur_uv[rc4][cc4]);   // Cb0 Cr0 Cb1 Cr1 at x = (lane&3)*2, +1
    const int x0 = (lane & 3) * 2;
    uint32_t qd, qh, qv, qp_ = 0;
    {
      const int blk = (rc4 >> 2) * 2 + (x0 >> 2);
      qd = (uint32_t)nb.cdc[blk] | ((uint32_t)nb.cdc[4 + blk] << 8);
      qd |= qd << 16;
      qh = (uint32_t)nb.left_uv[rc4 * 2] | ((uint32_t)nb.left_uv[rc4 * 2 + 1] << 8);
      qh |= qh << 16;
      qv = *reinterpret_cast<const uint32_t*>(&nb.top_uv[cc4]);
      int Hc[2] = {0, 0}, Vc[2] = {0, 0};
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int tlc = c ? nb.tl_v : nb.tl_u;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          Hc[c] += (i + 1) * ((int)nb.top_uv[(4 + i) * 2 + c] - (i == 3 ? tlc : (int)nb.top_uv[(2 - i) * 2 + c]));
          Vc[c] += (i + 1) * ((int)nb.left_uv[(4 + i) * 2 + c] - (i == 3 ? tlc : (int)nb.left_uv[(2 - i) * 2 + c]));
        }
        const int a = 16 * ((int)nb.left_uv[14 + c] + (int)nb.top_uv[14 + c]), b = (34 * Hc[c] + 32) >> 6, cq = (34 * Vc[c] + 32) >> 6;
        qp_ |= (uint32_t)clip255((a + b * (x0 - 3) + cq * (rc4 - 3) + 16) >> 5) << (8 * c);
        qp_ |= (uint32_t)clip255((a + b * (x0 + 1 - 3) + cq * (rc4 - 3) + 16) >> 5) << (16 + 8 * c);
      }
    }
    int cbest_key = 0x7fffffff, cbest = 0;
    {
      int s = __reduce_add_sync(FULL, __vsadu4(cc, qd)); if (s * 4 + 0 < cbest_key) { cbest_key = s * 4 + 0; cbest = 0; }
      if (has_left) { s = __reduce_add_sync(FULL, __vsadu4(cc, qh)); if (s * 4 + 1 < cbest_key) { cbest_key = s * 4 + 1; cbest = 1; } }
      if (has_top) { s = __reduce_add_sync(FULL, __vsadu4(cc, qv)); if (s * 4 + 2 < cbest_key) { cbest_key = s * 4 + 2; cbest = 2; } }
      if (has_top && has_left) { s = __reduce_add_sync(FULL, __vsadu4(cc, qp_)); if (s * 4 + 3 < cbest_key) { cbest_key = s * 4 + 3; cbest = 3; } }
    }
    *reinterpret_cast<uint32_t*>(&t.pred_uv[rc4][cc4]) = cbest == 0 ? qd : cbest == 1 ? qh : cbest == 2 ? qv : qp_;
    __syncwarp();
    // ---- transform / quantise / reconstruct ---------------------------------------------------------
    const int mb = mby * f.mbw + mbx;
    const int cbp = transform_mb<true>(t, lane, qp, f.coef + (size_t)mb * COEF_BLOCKS * 16, f.nnz + (size_t)mb * 32);
    __syncwarp();
    {
      const uint2 v = *reinterpret_cast<const uint2*>(&t.rec_y[r8][c8]);
      *reinterpret_cast<uint2*>(rec_y + (size_t)(mby * 16 + r8) * f.cw + mbx * 16 + c8) = v;
      if (lane < 16) {
        const uint2 w = *reinterpret_cast<const uint2*>(&t.rec_uv[r8][c8]);
        *reinterpret_cast<uint2*>(rec_uv + (size_t)(mby * 8 + r8) * f.cw + mbx * 16 + c8) = w;
      }
      // right column becomes the next macroblock's left neighbour
      if (lane < 16) nb.left_y[lane] = t.rec_y[lane][15];
      else nb.left_uv[lane - 16] = t.rec_uv[(lane - 16) >> 1][14 + ((lane - 16) & 1)];
      if (lane == 0) {
        MbInfo mi; mi.mvx = 0; mi.mvy = 0; mi.type = MB_I16; mi.i16_mode = (uint8_t)best_mode; mi.chroma_mode = (uint8_t)cbest; mi.cbp = (uint8_t)cbp;
        if (cbp < 0) { mi.type = MB_PCM; mi.cbp = 0; }   // too big for CAVLC: I_PCM (transform_mb)
        f.mbinfo[mb] = mi;
      }
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) *((volatile int*)&f.progress[mby]) = mbx + 1;
  }
}

int launch_intra(const FrameCtx& f, cudaStream_t st) {
  cudaMemsetAsync(f.progress, 0, sizeof(int) * f.mbh, st);
  k_intra_rows<<<f.mbh, 32, 0, st>>>(f);
  return 1;
}

}  // namespace b2v
