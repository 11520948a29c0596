// imp_blur.cuh — Gaussian blur (cvSmooth CV_GAUSSIAN, filters.c:192-207) fused with the op list and the
// oriented store, sm_100a.
//
// One CTA per 32 x 64 tile of the BASE frame, tiled in DESTINATION space (so that with any of the eight output
// orientations the tile's destination rows start 16-byte aligned and leave as 128-bit stores):
//   1. ONE TMA box load (cp.async.bulk.tensor.2d, SASS UTMALDG) of the (32+2R) x (64+2R) source neighbourhood, zero-filled
//      outside the tensor; border pixels are then replicated in shared memory (BORDER_REPLICATE).
//   2. planarise: interleaved B,G,R[,A] rows -> one byte plane per channel (PRMT byte transposes), so that four
//      horizontally adjacent samples of a channel share a 32-bit word.
//   3. horizontal pass with IDP.4A: the 8-bit taps are tabulated by the planner as four byte-shifted versions, so an output
//      at byte offset j of an aligned window is sum_w dp4a(W[j/4 + w], taps[j%4][w]): 4 instructions for 13 taps instead of
//      13 IMADs. Sums are exact u16 (<= 255*256) and are stored TRANSPOSED (row index contiguous).
//   4. vertical pass with IDP.2A (u16 x u8 pairs): two vertically adjacent sums share a word, taps in two parity versions:
//      R+1 instructions per output instead of 2R+1. (acc + 2^15) >> 16 is OpenCV's fixed-point rounding.
//      A thread owns one column of the tile and a run of 8 rows for ALL channels, so the blurred pixels stay in registers:
//   5. the op list runs on them right there, results are written to a shared-memory stage in destination orientation,
//   6. copy-out: 16 bytes per thread, coalesced rows.
// R is the tap radius padded to {3,6,9,12}; zero taps change nothing. Taps > 255 (a lone centre tap of 256) never get here.
#pragma once
#include "imp_tiles.cuh"

namespace imp_tiles {

constexpr int BTW = IMP_BLUR_TW, BTH = IMP_BLUR_TH;         // base-frame tile (imp_plan.h)
constexpr int BLUR_THREADS = 256;


template <int SC, int R, bool NOCOMP>
__global__ void __launch_bounds__(BLUR_THREADS, 4)
imp_blur_tile_kernel(const ImpJob* __restrict__ jobs, int first, int count, const __grid_constant__ ImpJob one) {
    using D = ImpBlurDims<R>;
    constexpr int SPANX = D::SPANX, SPANY = D::SPANY, NWH = D::NWH, NWIN = D::NWIN, PWW = D::PWW;
    constexpr int NDV = D::NDV, NWV = D::NWV, NV128 = D::NV128, HS = D::HS;
    extern __shared__ __align__(128) uint8_t smem[];
    const int jn = blockIdx.y + blockIdx.z * 65535;
    if (jn >= count) return;
    const ImpJob* __restrict__ jp = jobs ? jobs + first + jn : &one;
    ImpJob job;
    job.src = jp->src; job.dst = jp->dst; job.pass = jp->pass; job.wm = jp->wm;
    job.src_pitch = jp->src_pitch; job.dst_pitch = jp->dst_pitch; job.wm_pitch = jp->wm_pitch; job.wm_c = jp->wm_c; job.tm_x0 = jp->tm_x0;
    const uint8_t* __restrict__ blob = job.pass;
    const ImpPass* __restrict__ P = reinterpret_cast<const ImpPass*>(blob);
    const int w = P->sw, h = P->sh;                                     // blur: base frame == source window
    const ImpFrameMap om = P->out;
    // ---- this CTA's tile, chosen in destination space ----
    const int TWd = om.swap ? BTH : BTW, THd = om.swap ? BTW : BTH;
    const int xsh = om.swap ? 6 : 5, ysh = om.swap ? 5 : 6;             // log2 of the destination tile edges
    const int tiles_xd = (om.w + TWd - 1) >> xsh, tiles_yd = (om.h + THd - 1) >> ysh;
    if ((int)blockIdx.x >= tiles_xd * tiles_yd) return;
    const int X0 = ((int)blockIdx.x % tiles_xd) * TWd, Y0 = ((int)blockIdx.x / tiles_xd) * THd;
    const int vw = min(TWd, om.w - X0), vh = min(THd, om.h - Y0);       // valid destination rectangle
    const int ulo = om.flipx ? om.w - X0 - vw : X0, vlo = om.flipy ? om.h - Y0 - vh : Y0;
    const int x0 = om.swap ? vlo : ulo, y0 = om.swap ? ulo : vlo;       // base-frame origin of the tile

    const int rs = P->tile_rs;
    const int nops = P->nops;
    const int ops_bytes = (nops * (int)sizeof(ImpOp) + P->lut_bytes + 15) & ~15;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint8_t* s_ops = smem + 128;
    uint8_t* tile = s_ops + ((ops_bytes + 127) & ~127);                 // TMA box; later the destination-oriented out stage
    uint32_t* planar = reinterpret_cast<uint32_t*>(tile + ((rs * SPANY + 127) & ~127));     // [SC][SPANY][PWW] words; later the blurred-byte stage
    uint16_t* hbuf = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(planar) + ((SC * SPANY * PWW * 4 + 127) & ~127));   // [SC*BTW][HS]
    uint8_t* ostage = tile;
    const int tid = threadIdx.x;
    // box origin: source pixel (x0-R, y0-R); 16-byte aligned in x as TMA requires (coordinates may be negative)
    const int xbyte = job.tm_x0 + (x0 - R) * SC;
    const int c0 = (xbyte >> 4) << 1;
    const int col_off = job.tm_x0 - c0 * 8;                             // tile byte offset of source pixel 0
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)(rs * SPANY));
        tma_load_2d(tile, jp->tmap, c0, y0 - R, bar);
    }
    {
        const uint4* gsrc = reinterpret_cast<const uint4*>(blob + P->ops_off);
        uint4* sdst = reinterpret_cast<uint4*>(s_ops);
        for (int i = tid; i < ops_bytes / 16; i += BLUR_THREADS) sdst[i] = __ldg(gsrc + i);
    }
    __syncthreads();
    mbar_wait(bar, 0);

    // BORDER_REPLICATE: fill the part of the neighbourhood that lies outside the image from the clamped pixel
    if (x0 - R < 0 || y0 - R < 0 || x0 + BTW + R > w || y0 + BTH + R > h) {
        for (int i = tid; i < SPANX * SPANY; i += BLUR_THREADS) {
            const int ty = i / SPANX, tx = i - ty * SPANX;
            const int X = x0 - R + tx, Y = y0 - R + ty;
            const int cx = min(max(X, 0), w - 1), cy = min(max(Y, 0), h - 1);
            if (cx != X || cy != Y) {
                const uint8_t* s = tile + (cy - (y0 - R)) * rs + col_off + cx * SC;
                uint8_t* d = tile + ty * rs + col_off + X * SC;
                IMP_DBG(s >= tile && s + SC <= tile + rs * SPANY && d >= tile && d + SC <= tile + rs * SPANY, 0);
#pragma unroll
                for (int c = 0; c < SC; c++) d[c] = s[c];
            }
        }
        __syncthreads();
    }

    // ---- planarise: item = (row, group of 4 pixels) ----
    {
        constexpr int NG = (SPANX + 3) / 4;
        const uint8_t* trow0 = tile + col_off + (x0 - R) * SC;
        for (int item = tid; item < SPANY * NG; item += BLUR_THREADS) {
            const int row = item / NG, g = item - row * NG;
            const uint8_t* p = trow0 + row * rs + g * 4 * SC;
            uint32_t* dst = planar + row * PWW + g;
            if (SC == 4) {
                const uint32_t* pw = reinterpret_cast<const uint32_t*>(p);
                const uint32_t w0 = pw[0], w1 = pw[1], w2 = pw[2], w3 = pw[3];
                const uint32_t t0 = __byte_perm(w0, w1, 0x5140), t1 = __byte_perm(w2, w3, 0x5140);
                const uint32_t t2 = __byte_perm(w0, w1, 0x7362), t3 = __byte_perm(w2, w3, 0x7362);
                dst[0] = __byte_perm(t0, t1, 0x5410);
                dst[(SC > 1 ? 1 : 0) * SPANY * PWW] = __byte_perm(t0, t1, 0x7632);
                dst[(SC > 2 ? 2 : 0) * SPANY * PWW] = __byte_perm(t2, t3, 0x5410);
                dst[(SC > 3 ? 3 : 0) * SPANY * PWW] = __byte_perm(t2, t3, 0x7632);
            } else {
                uint32_t wv[3];
                load_bytes<12>(p, wv);
                dst[0] = __byte_perm(__byte_perm(wv[0], wv[1], 0x0630), wv[2], 0x5210);
                dst[(SC > 1 ? 1 : 0) * SPANY * PWW] = __byte_perm(__byte_perm(wv[0], wv[1], 0x0741), wv[2], 0x6210);
                dst[(SC > 2 ? 2 : 0) * SPANY * PWW] = __byte_perm(__byte_perm(wv[0], wv[1], 0x0052), wv[2], 0x7410);
            }
        }
    }
    __syncthreads();

    // ---- horizontal pass (IDP.4A): item = (channel, 8-pixel group, row), rows fastest ----
    {
        uint32_t th[4][NWH];
        const uint32_t* taph = reinterpret_cast<const uint32_t*>(blob + P->taph_off);
#pragma unroll
        for (int m = 0; m < 4; m++)
#pragma unroll
            for (int i = 0; i < NWH; i++) th[m][i] = __ldg(taph + m * NWH + i);
        for (int item = tid; item < SC * 4 * SPANY; item += BLUR_THREADS) {
            const int row = item % SPANY, rest = item / SPANY, g = rest & 3, c = rest >> 2;
            const uint32_t* pw = planar + (c * SPANY + row) * PWW + 2 * g;
            uint32_t W[NWIN];
#pragma unroll
            for (int i = 0; i < NWIN; i++) W[i] = pw[i];
            uint16_t* hp = hbuf + (c * BTW + 8 * g) * HS + row;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                unsigned acc = 0;
#pragma unroll
                for (int i = 0; i < NWH; i++) acc = __dp4a(W[(j >> 2) + i], th[j & 3][i], acc);
                hp[j * HS] = (uint16_t)acc;
            }
        }
    }
    __syncthreads();

    // ---- vertical pass (IDP.2A) fused with the op list: a thread owns column x of the tile and a run of 8 rows, all
    // channels, so the blurred pixels never leave registers before the ops have run; results go to the out stage in
    // destination orientation ----
    static_assert(BTW == 32 && BTH == 64 && BLUR_THREADS == BTW * (BTH / 8), "one (column, 8-row run) per thread");
    const int oc = P->oc, dc = P->dc;
    // out-stage row stride: an ODD number of words, so that the 32 lanes of a warp — which own 32 different destination rows
    // when the output is rotated — land in 32 different banks (a stride of 48 or 64 words put them into two)
    const int OS = TWd * dc + 4;
    const int tw = om.swap ? vh : vw, th = om.swap ? vw : vh;           // base-frame extent of the tile
    {
        uint32_t tv[2][NWV];
        const uint32_t* tapv = reinterpret_cast<const uint32_t*>(blob + P->tapv_off);
#pragma unroll
        for (int m = 0; m < 2; m++)
#pragma unroll
            for (int i = 0; i < NWV; i++) tv[m][i] = __ldg(tapv + m * NWV + i);
        const int x = tid & 31, g = tid >> 5;
        uint32_t lo[SC], hi[SC];                                        // blurred bytes of rows 0-3 / 4-7 of the run, per channel
#pragma unroll
        for (int c = 0; c < SC; c++) {
            const uint4* hp = reinterpret_cast<const uint4*>(hbuf + (c * BTW + x) * HS + 8 * g);
            uint32_t W[4 * NV128];
#pragma unroll
            for (int i = 0; i < NV128; i++) { const uint4 q = hp[i]; W[4 * i] = q.x; W[4 * i + 1] = q.y; W[4 * i + 2] = q.z; W[4 * i + 3] = q.w; }
            uint32_t r[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                unsigned acc = 32768u;
#pragma unroll
                for (int t = 0; t < NDV; t++) {
                    const uint32_t a = W[(j >> 1) + t], b = tv[j & 1][t >> 1];
                    acc = (t & 1) ? __dp2a_hi(a, b, acc) : __dp2a_lo(a, b, acc);
                }
                r[j] = acc;                                             // the blurred byte is bits 16..23 (sum of taps = 256)
            }
            lo[c] = __byte_perm(__byte_perm(r[0], r[1], 0x0062), __byte_perm(r[2], r[3], 0x0062), 0x5410);
            hi[c] = __byte_perm(__byte_perm(r[4], r[5], 0x0062), __byte_perm(r[6], r[7], 0x0062), 0x5410);
        }
        const int bx = x0 + x;
        // pixels of a partial tile beyond the frame (up to 31 columns / 63 rows, < IMP_VIGNETTE_MARGIN) run the ops on their own
        // out-of-frame coordinates and are never stored: position-dependent ops only compare coordinates or index tables
        // that carry that margin
        // the out-stage address of base pixel (bx, by) is affine in by for this thread's column (cf. StripStore)
        int so0, sstep;
        {
            int Xa, Ya, Xb, Yb;
            imp_map_xy(om, bx, y0, Xa, Ya);
            imp_map_xy(om, bx, y0 + 1, Xb, Yb);
            so0 = (Ya - Y0) * OS + (Xa - X0) * dc;
            sstep = (Yb - Ya) * OS + (Xb - Xa) * dc;
        }
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            ImpPx px[4];
            int bxs[4], bys[4];
            const uint32_t wb = half ? hi[0] : lo[0], wg = half ? hi[SC > 1 ? 1 : 0] : lo[SC > 1 ? 1 : 0];
            const uint32_t wr = half ? hi[SC > 2 ? 2 : 0] : lo[SC > 2 ? 2 : 0], wa = half ? hi[SC - 1] : lo[SC - 1];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                // byte k of the packed rows, zero-extended: one PRMT instead of a shift and a mask
                px[k].b = (int)__byte_perm(wb, 0u, 0x4440u + k);
                px[k].g = (int)__byte_perm(wg, 0u, 0x4440u + k);
                px[k].r = (int)__byte_perm(wr, 0u, 0x4440u + k);
                px[k].a = (SC == 4) ? (int)__byte_perm(wa, 0u, 0x4440u + k) : 255;
                bxs[k] = bx; bys[k] = y0 + 8 * g + 4 * half + k;      // unclamped (see below): affine in k, one add per pixel in the ops
            }
            if (nops) imp_run_ops_n<4, false, NOCOMP>(px, oc, bxs, bys, reinterpret_cast<const ImpOp*>(s_ops), nops, s_ops + nops * sizeof(ImpOp), job.wm, job.wm_pitch, job.wm_c);
            const int ly0 = 8 * g + 4 * half;
            if (x >= tw || ly0 >= th) continue;                         // outside the tile's valid rectangle (frame edge)
            const int nk = min(4, th - ly0);                            // 4 except in the last rows of a frame
            uint8_t* d = ostage + so0 + ly0 * sstep;
            IMP_DBG(d >= ostage && d + dc <= ostage + OS * THd && d + (nk - 1) * sstep >= ostage && d + (nk - 1) * sstep + dc <= ostage + OS * THd, 1);
            if (dc == 4) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k < nk) *reinterpret_cast<uchar4*>(d + k * sstep) = make_uchar4((unsigned char)px[k].b, (unsigned char)px[k].g, (unsigned char)px[k].r, (unsigned char)px[k].a);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k < nk) { uint8_t* e = d + k * sstep; e[0] = (unsigned char)px[k].b; e[1] = (unsigned char)px[k].g; e[2] = (unsigned char)px[k].r; }
                }
            }
        }
    }
    __syncthreads();
    // ---- copy-out: rows of the destination rectangle, 16 bytes per thread when the rows are 16-byte addressable ----
    tile_copy_out_w(ostage, OS, job.dst + (size_t)Y0 * job.dst_pitch + (size_t)X0 * dc, job.dst_pitch, vw * dc, vh, tid, BLUR_THREADS);
}

}  // namespace imp_tiles
