// imp_gathertile.cuh — the plain gathers (index map / crop copy, INTER_NN, INTER_LINEAR; bridge.c:130-137, 190-191,
// SURVEY App. A.1/A.4) from a TMA-staged source tile, sm_100a.
//
// One CTA per T x T tile (T = 64 or 32, chosen by the planner from the source footprint) of the DESTINATION image, so the
// tile is a square of the base frame under all eight output orientations and its destination rows are contiguous:
//   1. ONE TMA box load (cp.async.bulk.tensor.2d, SASS UTMALDG) of the source rectangle the tile reads;
//   2. the tile's column and row lookups (source offsets, bilinear coefficients) are resolved once into shared memory;
//   3. a thread owns FOUR horizontally adjacent destination pixels per step: four independent gathers, the op list runs
//      once over the four (imp_run_ops_n<4>: parameter fetch and dispatch are shared), and 4-channel results leave as ONE
//      128-bit store per thread (512 contiguous bytes per warp); 3-channel results are assembled in a shared-memory stage
//      and copied out as 16-byte chunks.
// Round 1 ran these gathers one pixel per thread straight from global memory with byte stores for 3 channels; the strip
// kernel variants of the first half of round 2 staged the source but still paid every per-row cost per pixel.
#pragma once
#include "imp_blur.cuh"      // tile_copy_out, IMP_DBG

namespace imp_tiles {

constexpr int GATHER_THREADS = 256;
// (IMP_GATHER_TPC, imp_plan.h: tiles a CTA walks)

// shared-memory words the lookups take: per column and per row {offset 0, offset 1, coefficients}
__host__ __device__ constexpr int gather_table_words(int T) { return 6 * T; }

template <int SC, int KIND, bool LIGHT>
__global__ void __launch_bounds__(GATHER_THREADS, LIGHT ? 4 : 3)
imp_gather_tile_kernel(const ImpJob* __restrict__ jobs, int first, int count, const __grid_constant__ ImpJob one) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int jn = blockIdx.y + blockIdx.z * 65535;
    if (jn >= count) return;
    const ImpJob* __restrict__ jp = jobs ? jobs + first + jn : &one;
    ImpJob job;
    job.src = jp->src; job.dst = jp->dst; job.pass = jp->pass; job.wm = jp->wm;
    job.src_pitch = jp->src_pitch; job.dst_pitch = jp->dst_pitch; job.wm_pitch = jp->wm_pitch; job.wm_c = jp->wm_c; job.tm_x0 = jp->tm_x0;
    const uint8_t* __restrict__ blob = job.pass;
    const ImpPass* __restrict__ P = reinterpret_cast<const ImpPass*>(blob);
    const int T = P->gt;
    const int sw = P->sw, sh = P->sh;
    const ImpFrameMap om = P->out;
    const int tsh = T == 64 ? 6 : 5;                                    // T is 64 or 32
    const int tiles_xd = (om.w + T - 1) >> tsh, tiles_yd = (om.h + T - 1) >> tsh;
    // a CTA walks IMP_GATHER_TPC consecutive tiles: the job / pass / op-list prologue is paid once, and the TMA box of tile
    // i+1 is in flight while tile i is processed (two boxes, two mbarriers)
    const int t_begin = (int)blockIdx.x * IMP_GATHER_TPC, t_end = min(t_begin + IMP_GATHER_TPC, tiles_xd * tiles_yd);
    if (t_begin >= t_end) return;

    const int* __restrict__ xofs = reinterpret_cast<const int*>(blob + P->xofs_off);
    const int* __restrict__ yofs = reinterpret_cast<const int*>(blob + P->yofs_off);
    // clamped source coordinate(s) of a base column / row
    auto xs0 = [&](int bx) { return KIND == IMP_G_COPY ? bx : KIND == IMP_G_NN ? __ldg(xofs + bx) : min(max(__ldg(xofs + bx), 0), sw - 1); };
    auto xs1 = [&](int bx) { return min(max(__ldg(xofs + bx) + 1, 0), sw - 1); };
    auto ys0 = [&](int by) { return KIND == IMP_G_COPY ? by : KIND == IMP_G_NN ? __ldg(yofs + by) : min(max(__ldg(yofs + by), 0), sh - 1); };
    auto ys1 = [&](int by) { return min(max(__ldg(yofs + by) + 1, 0), sh - 1); };

    const int rs = P->tile_rs;
    const int nops = P->nops;
    const int ops_bytes = (nops * (int)sizeof(ImpOp) + P->lut_bytes + 15) & ~15;
    const int box_bytes = (rs * P->tile_rows + 127) & ~127;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);                  // [2]
    uint8_t* s_ops = smem + 128;
    uint8_t* box0 = s_ops + ((ops_bytes + 127) & ~127);                 // two TMA boxes
    int* tab = reinterpret_cast<int*>(box0 + 2 * box_bytes);
    int* cx0 = tab, * cx1 = tab + T, * cxa = tab + 2 * T, * ry0 = tab + 3 * T, * ry1 = tab + 4 * T, * ryb = tab + 5 * T;
    uint8_t* ostage = reinterpret_cast<uint8_t*>(tab + gather_table_words(T));      // T rows of T*3 bytes (3-channel results)
    const int tid = threadIdx.x;
    const int oc = P->oc, dc = P->dc;
    const int OS = T * 3;                                               // out-stage row stride (3-channel results)
    const int qsh = tsh - 2;                                            // log2 of the 4-pixel groups per tile row
    const bool vec16 = dc == 4 && ((reinterpret_cast<uintptr_t>(job.dst) | (unsigned)job.dst_pitch) & 15) == 0;

    // geometry of tile t: destination rectangle, base-frame origin and extent, first source pixel / row of its box
    struct Geo { int X0, Y0, vw, vh, x0, y0, tw, th, sx_first, sy_first; };
    auto geometry = [&](int t) {
        Geo g;
        const int ty = t / tiles_xd;
        g.X0 = (t - ty * tiles_xd) << tsh; g.Y0 = ty << tsh;
        g.vw = min(T, om.w - g.X0); g.vh = min(T, om.h - g.Y0);         // valid destination rectangle
        const int ulo = om.flipx ? om.w - g.X0 - g.vw : g.X0, vlo = om.flipy ? om.h - g.Y0 - g.vh : g.Y0;
        g.x0 = om.swap ? vlo : ulo; g.y0 = om.swap ? ulo : vlo;         // base-frame origin of the tile
        g.tw = om.swap ? g.vh : g.vw; g.th = om.swap ? g.vw : g.vh;     // base-frame extent of the tile
        g.sx_first = xs0(g.x0); g.sy_first = ys0(g.y0);                 // the tables are non-decreasing
        return g;
    };
    auto issue = [&](const Geo& g, int buf) {                           // one thread: box of tile g into buffer buf
        const int xbyte = job.tm_x0 + g.sx_first * SC;
        mbar_expect_tx(bar + buf, (uint32_t)(rs * P->tile_rows));
        tma_load_2d(box0 + buf * box_bytes, jp->tmap, (xbyte >> 4) << 1, g.sy_first, bar + buf);
    };
    Geo g = geometry(t_begin);
    if (tid == 0) {
        mbar_init(bar, 1); mbar_init(bar + 1, 1);
        issue(g, 0);
    }
    {
        const uint4* gsrc = reinterpret_cast<const uint4*>(blob + P->ops_off);
        uint4* sdst = reinterpret_cast<uint4*>(s_ops);
        for (int i = tid; i < ops_bytes / 16; i += GATHER_THREADS) sdst[i] = __ldg(gsrc + i);
    }
    for (int t = t_begin, it = 0; t < t_end; t++, it++) {
    const int buf = it & 1;
    const uint8_t* tile = box0 + buf * box_bytes;
    const int X0 = g.X0, Y0 = g.Y0, vw = g.vw, vh = g.vh, x0 = g.x0, y0 = g.y0, tw = g.tw, th = g.th, sx_first = g.sx_first, sy_first = g.sy_first;
    const int xbyte = job.tm_x0 + sx_first * SC;
    const int col_off = xbyte - ((xbyte >> 4) << 4);                    // tile byte offset of source pixel sx_first
    if (it > 0) __syncthreads();                                        // the previous tile's lookups (and its box) are no longer read
    if (t + 1 < t_end) {
        g = geometry(t + 1);
        if (tid == 0) { fence_generic_to_async_smem(); issue(g, buf ^ 1); }     // that buffer was read two tiles ago, behind the barrier above
    }
    // the tile's lookups, resolved once: byte offset of each base column inside a tile row, of each base row inside the tile
    if (tid < tw) {
        const int bx = x0 + tid;
        cx0[tid] = col_off + (xs0(bx) - sx_first) * SC;
        if (KIND == IMP_G_LINEAR) {
            cx1[tid] = col_off + (xs1(bx) - sx_first) * SC;
            cxa[tid] = __ldg(reinterpret_cast<const int*>(blob + P->xcoef_off) + bx);           // two shorts
            IMP_DBG(cx1[tid] >= 0 && cx1[tid] + SC <= rs, 4);
        }
        IMP_DBG(cx0[tid] >= 0 && cx0[tid] + SC <= rs, 4);
    } else if (tid >= 64 && tid - 64 < th) {
        const int i = tid - 64, by = y0 + i;
        ry0[i] = (ys0(by) - sy_first) * rs;
        if (KIND == IMP_G_LINEAR) {
            ry1[i] = (ys1(by) - sy_first) * rs;
            ryb[i] = __ldg(reinterpret_cast<const int*>(blob + P->ycoef_off) + by);
            IMP_DBG(ry1[i] >= 0 && ry1[i] < rs * P->tile_rows, 5);
        }
        IMP_DBG(ry0[i] >= 0 && ry0[i] < rs * P->tile_rows, 5);
    }
    __syncthreads();
    mbar_wait(bar + buf, (it >> 1) & 1);

    const int dstep = om.flipx ? -1 : 1, dbx = om.swap ? 0 : dstep, dby = om.swap ? dstep : 0;
    for (int item = tid; item < (T << qsh); item += GATHER_THREADS) {
        const int Yl = item >> qsh, Xq = (item & ((1 << qsh) - 1)) * 4;
        if (Yl >= vh || Xq >= vw) continue;
        ImpPx px[4];
        int bxs[4], bys[4];
        // base-frame coordinates of the group's first pixel; the next ones are one step along the base x or y axis
        // (destination x runs along base x, or along base y when the output is transposed; backwards when flipped)
        int bx0, by0;
        {
            const int X = X0 + Xq, Y = Y0 + Yl;
            const int u = om.flipx ? om.w - 1 - X : X, v = om.flipy ? om.h - 1 - Y : Y;
            bx0 = om.swap ? v : u; by0 = om.swap ? u : v;
        }
        const int kmax = vw - 1 - Xq;                                   // dead slots (k > kmax) compute on the last valid pixel, never stored
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int kk = min(k, kmax);
            const int bx = bx0 + kk * dbx, by = by0 + kk * dby;
            bxs[k] = bx; bys[k] = by;
            const int lx = bx - x0, ly = by - y0;
            IMP_DBG(lx >= 0 && lx < tw && ly >= 0 && ly < th, 6);
            int v4[4] = {0, 0, 0, 255};
            if (KIND == IMP_G_LINEAR) {
                const uint8_t* q0 = tile + ry0[ly]; const uint8_t* q1 = tile + ry1[ly];
                const int o0 = cx0[lx], o1 = cx1[lx];
                const int av = cxa[lx], bv = ryb[ly];
                const int a0 = (short)(av & 0xffff), a1 = av >> 16, b0 = (short)(bv & 0xffff), b1 = bv >> 16;
                int p00[SC], p01[SC], p10[SC], p11[SC];
                if (SC == 4) {
                    const uint32_t w00 = *reinterpret_cast<const uint32_t*>(q0 + o0), w01 = *reinterpret_cast<const uint32_t*>(q0 + o1);
                    const uint32_t w10 = *reinterpret_cast<const uint32_t*>(q1 + o0), w11 = *reinterpret_cast<const uint32_t*>(q1 + o1);
#pragma unroll
                    for (int c = 0; c < SC; c++) { p00[c] = (w00 >> (8 * c)) & 255; p01[c] = (w01 >> (8 * c)) & 255; p10[c] = (w10 >> (8 * c)) & 255; p11[c] = (w11 >> (8 * c)) & 255; }
                } else {
#pragma unroll
                    for (int c = 0; c < SC; c++) { p00[c] = q0[o0 + c]; p01[c] = q0[o1 + c]; p10[c] = q1[o0 + c]; p11[c] = q1[o1 + c]; }
                }
#pragma unroll
                for (int c = 0; c < SC; c++) {                          // SURVEY App. A.4 / imp_gather_linear
                    const int h0 = p00[c] * a0 + p01[c] * a1, h1 = p10[c] * a0 + p11[c] * a1;
                    v4[c] = ((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2) & 255;
                }
            } else {
                const uint8_t* q = tile + ry0[ly] + cx0[lx];
                if (SC == 4) {
                    const uint32_t w = *reinterpret_cast<const uint32_t*>(q);
                    v4[0] = __byte_perm(w, 0, 0x4440); v4[1] = __byte_perm(w, 0, 0x4441); v4[2] = __byte_perm(w, 0, 0x4442); v4[3] = w >> 24;
                } else {
#pragma unroll
                    for (int c = 0; c < SC; c++) v4[c] = q[c];
                }
            }
            if (SC == 1) { px[k].b = px[k].g = px[k].r = v4[0]; px[k].a = 255; }
            else { px[k].b = v4[0]; px[k].g = v4[1]; px[k].r = v4[2]; px[k].a = (SC == 4) ? v4[3] : 255; }
        }
        if (nops) imp_run_ops_n<4, LIGHT>(px, oc, bxs, bys, reinterpret_cast<const ImpOp*>(s_ops), nops, s_ops + nops * sizeof(ImpOp), job.wm, job.wm_pitch, job.wm_c);
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; k++)         // the low bytes of b, g, r, a: three PRMTs
            w[k] = __byte_perm(__byte_perm((unsigned)px[k].b, (unsigned)px[k].g, 0x0040), __byte_perm((unsigned)px[k].r, (unsigned)px[k].a, 0x0040), 0x5410);
        if (dc == 4) {
            uint8_t* d = job.dst + (size_t)(Y0 + Yl) * job.dst_pitch + (size_t)(X0 + Xq) * 4;
            if (vec16 && Xq + 4 <= vw) *reinterpret_cast<uint4*>(d) = make_uint4(w[0], w[1], w[2], w[3]);
            else {
#pragma unroll
                for (int k = 0; k < 4; k++) if (Xq + k < vw) reinterpret_cast<uint32_t*>(d)[k] = w[k];
            }
        } else {
            // 12 bytes = three aligned words of the out stage (row stride T*3 and Xq*3 are multiples of 4); the bytes past vw
            // of the last group stay inside the stage row and are not copied out
            uint32_t* d = reinterpret_cast<uint32_t*>(ostage + Yl * OS + Xq * 3);
            d[0] = (w[0] & 0xFFFFFFu) | (w[1] << 24);
            d[1] = ((w[1] >> 8) & 0xFFFFu) | (w[2] << 16);
            d[2] = ((w[2] >> 16) & 0xFFu) | (w[3] << 8);
        }
    }
    if (dc != 4) {
        __syncthreads();
        tile_copy_out(ostage, OS, job.dst + (size_t)Y0 * job.dst_pitch + (size_t)X0 * 3, job.dst_pitch, vw * 3, vh, tid, GATHER_THREADS);
    }
    }   // tiles of this CTA
}

}  // namespace imp_tiles
