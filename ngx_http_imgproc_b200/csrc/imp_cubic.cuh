// imp_cubic.cuh — INTER_CUBIC (cvResize, bridge.c:190-191) from a TMA-staged source tile, sm_100a.
//
// One CTA per 32 x 32 output tile, tiled in DESTINATION space (aligned 16-byte rows under every output orientation):
//   1. ONE TMA box load of the source rectangle the tile's 4x4 footprints touch (clamped to the window: OpenCV
//      replicates the border by clamping tap coordinates, so no fill is needed);
//   2. horizontal pass ONCE per (source row, output column): OpenCV's int32 H = sum p*a (11-bit coefficients), kept as
//      exact floats in shared memory (|H| < 2^22). At 2x that is ~0.6 horizontal rows per output pixel; the per-pixel
//      gather pays 4, the column-run kernel of round 1 paid 0.94;
//   3. vertical pass per output pixel from shared memory in OpenCV's float form (mul and add rounded separately, in its
//      order), bytes of a row at or beyond simd_end in the integer form of the scalar tail;
//   4. op list in registers, then 4-channel results leave as one 32-bit store per lane (128 contiguous bytes per warp),
//      3-channel results through a shared-memory stage as 16-byte stores.
#pragma once
#include "imp_blur.cuh"

namespace imp_tiles {

constexpr int CT = IMP_CUBIC_T;                 // output tile edge (imp_plan.h)
constexpr int CUBIC_THREADS = 256;

template <int SC> struct CubicH { static constexpr int HRS = IMP_CUBIC_HRS(SC); };   // floats per hbuf row (bank spread)

template <int SC>
__global__ void __launch_bounds__(CUBIC_THREADS, 4)
imp_cubic_tile_kernel(const ImpJob* __restrict__ jobs, int first, int count, const __grid_constant__ ImpJob one) {
    constexpr int HRS = CubicH<SC>::HRS;
    extern __shared__ __align__(128) uint8_t smem[];
    const int jn = blockIdx.y + blockIdx.z * 65535;
    if (jn >= count) return;
    const ImpJob* __restrict__ jp = jobs ? jobs + first + jn : &one;
    ImpJob job;
    job.src = jp->src; job.dst = jp->dst; job.pass = jp->pass; job.wm = jp->wm;
    job.src_pitch = jp->src_pitch; job.dst_pitch = jp->dst_pitch; job.wm_pitch = jp->wm_pitch; job.wm_c = jp->wm_c; job.tm_x0 = jp->tm_x0;
    const uint8_t* __restrict__ blob = job.pass;
    const ImpPass* __restrict__ P = reinterpret_cast<const ImpPass*>(blob);
    const int bw = P->bw, bh = P->bh, sw = P->sw, sh = P->sh;
    const ImpFrameMap om = P->out;
    const int tiles_xd = (om.w + CT - 1) / CT, tiles_yd = (om.h + CT - 1) / CT;
    if ((int)blockIdx.x >= tiles_xd * tiles_yd) return;
    const int X0 = ((int)blockIdx.x % tiles_xd) * CT, Y0 = ((int)blockIdx.x / tiles_xd) * CT;
    const int vw = min(CT, om.w - X0), vh = min(CT, om.h - Y0);         // valid destination rectangle
    const int ulo = om.flipx ? om.w - X0 - vw : X0, vlo = om.flipy ? om.h - Y0 - vh : Y0;
    const int x0 = om.swap ? vlo : ulo, y0 = om.swap ? ulo : vlo;       // base-frame origin of the tile
    const int tw = om.swap ? vh : vw, th = om.swap ? vw : vh;           // base-frame extent of the tile

    const int* __restrict__ xofs = reinterpret_cast<const int*>(blob + P->xofs_off);
    const short* __restrict__ xa = reinterpret_cast<const short*>(blob + P->xcoef_off);
    const int* __restrict__ yofs = reinterpret_cast<const int*>(blob + P->yofs_off);
    // source rectangle of the tile (clamped tap coordinates)
    const int sx_first = min(max(__ldg(xofs + x0) - 1, 0), sw - 1);
    const int sy_first = min(max(__ldg(yofs + y0) - 1, 0), sh - 1);
    const int sy_last = min(max(__ldg(yofs + y0 + th - 1) + 2, 0), sh - 1);
    const int nrows = sy_last - sy_first + 1;

    const int rs = P->tile_rs;
    const int nops = P->nops;
    const int ops_bytes = (nops * (int)sizeof(ImpOp) + P->lut_bytes + 15) & ~15;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint8_t* s_ops = smem + 128;
    uint8_t* tile = s_ops + ((ops_bytes + 127) & ~127);                 // TMA box; later the out stage of 3-channel results
    float* hbuf = reinterpret_cast<float*>(tile + IMP_CUBIC_TILE_BYTES(rs, P->tile_rows));  // [tile_rows][HRS]
    const int tid = threadIdx.x;
    const int xbyte = job.tm_x0 + sx_first * SC;
    const int c0 = (xbyte >> 4) << 1;
    const int col_off = xbyte - c0 * 8;                                 // tile byte offset of source pixel sx_first
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)(rs * P->tile_rows));
        tma_load_2d(tile, jp->tmap, c0, sy_first, bar);
    }
    {
        const uint4* gsrc = reinterpret_cast<const uint4*>(blob + P->ops_off);
        uint4* sdst = reinterpret_cast<uint4*>(s_ops);
        for (int i = tid; i < ops_bytes / 16; i += CUBIC_THREADS) sdst[i] = __ldg(gsrc + i);
    }
    // ---- horizontal pass: a thread owns one output column of the tile (lane) and walks the source rows ----
    {
        const int lx = tid & 31;
        const int bx = min(x0 + lx, bw - 1);
        const int xo = __ldg(xofs + bx) - 1;
        int so[4];                                                      // byte offsets of the four taps inside a tile row
#pragma unroll
        for (int t = 0; t < 4; t++) so[t] = col_off + (min(max(xo + t, 0), sw - 1) - sx_first) * SC;
        const int2 av = __ldg(reinterpret_cast<const int2*>(xa + bx * 4));   // 4 shorts, 8-byte aligned
        const int a0 = (short)(av.x & 0xffff), a1 = av.x >> 16, a2 = (short)(av.y & 0xffff), a3 = av.y >> 16;
        IMP_DBG(so[0] >= 0 && so[3] + SC <= rs && nrows <= P->tile_rows, 2);
        __syncthreads();
        mbar_wait(bar, 0);
        for (int r = tid >> 5; r < nrows; r += CUBIC_THREADS / 32) {
            const uint8_t* row = tile + r * rs;
            int hsum[SC];
            if (SC == 4) {
                const uint32_t p0 = *reinterpret_cast<const uint32_t*>(row + so[0]), p1 = *reinterpret_cast<const uint32_t*>(row + so[1]);
                const uint32_t p2 = *reinterpret_cast<const uint32_t*>(row + so[2]), p3 = *reinterpret_cast<const uint32_t*>(row + so[3]);
#pragma unroll
                for (int c = 0; c < SC; c++)
                    hsum[c] = (int)((p0 >> (8 * c)) & 255) * a0 + (int)((p1 >> (8 * c)) & 255) * a1 + (int)((p2 >> (8 * c)) & 255) * a2 + (int)((p3 >> (8 * c)) & 255) * a3;
            } else {
#pragma unroll
                for (int c = 0; c < SC; c++)
                    hsum[c] = (int)row[so[0] + c] * a0 + (int)row[so[1] + c] * a1 + (int)row[so[2] + c] * a2 + (int)row[so[3] + c] * a3;
            }
            float* hp = hbuf + r * HRS + lx * SC;
            if (SC == 4) *reinterpret_cast<float4*>(hp) = make_float4(imp_i2f22(hsum[0]), imp_i2f22(hsum[SC > 1 ? 1 : 0]), imp_i2f22(hsum[SC > 2 ? 2 : 0]), imp_i2f22(hsum[SC > 3 ? 3 : 0]));
            else {
#pragma unroll
                for (int c = 0; c < SC; c++) hp[c] = imp_i2f22(hsum[c]);
            }
        }
    }
    __syncthreads();

    // ---- vertical pass + op list: a thread owns 4 destination pixels of a column of the destination tile ----
    const short* __restrict__ yb = reinterpret_cast<const short*>(blob + P->ycoef_off);
    const float4* __restrict__ ybf = reinterpret_cast<const float4*>(blob + P->taps_off);
    const int oc = P->oc, dc = P->dc, simd_end = P->simd_end;
    const int OS = CT * 3;                                              // out-stage row stride (3-channel results), 96
    uint8_t* ostage = tile;
    ImpPx px[4];
    int bxs[4], bys[4];
    bool live[4];
#pragma unroll
    for (int it = 0; it < 4; it++) {
        int Xl = tid & 31, Yl = (tid >> 5) + it * 8;
        live[it] = Xl < vw && Yl < vh;
        Xl = min(Xl, vw - 1); Yl = min(Yl, vh - 1);                     // dead slots compute on a valid pixel, never stored
        const int X = X0 + Xl, Y = Y0 + Yl;
        const int u = om.flipx ? om.w - 1 - X : X, v = om.flipy ? om.h - 1 - Y : Y;
        const int bx = om.swap ? v : u, by = om.swap ? u : v;
        bxs[it] = bx; bys[it] = by;
        const int yo = __ldg(yofs + by) - 1;
        const float* h0 = hbuf + (min(max(yo, 0), sh - 1) - sy_first) * HRS + (bx - x0) * SC;
        const float* h1 = hbuf + (min(max(yo + 1, 0), sh - 1) - sy_first) * HRS + (bx - x0) * SC;
        const float* h2 = hbuf + (min(max(yo + 2, 0), sh - 1) - sy_first) * HRS + (bx - x0) * SC;
        const float* h3 = hbuf + (min(max(yo + 3, 0), sh - 1) - sy_first) * HRS + (bx - x0) * SC;
        IMP_DBG(h0 >= hbuf && h3 + SC <= hbuf + P->tile_rows * HRS && h3 >= hbuf && h0 + SC <= hbuf + P->tile_rows * HRS && bx - x0 >= 0 && bx - x0 < CT, 3);
        float f0[SC], f1[SC], f2[SC], f3[SC];
        if (SC == 4) {
            const float4 q0 = *reinterpret_cast<const float4*>(h0), q1 = *reinterpret_cast<const float4*>(h1);
            const float4 q2 = *reinterpret_cast<const float4*>(h2), q3 = *reinterpret_cast<const float4*>(h3);
            f0[0] = q0.x; f0[SC > 1 ? 1 : 0] = q0.y; f0[SC > 2 ? 2 : 0] = q0.z; f0[SC > 3 ? 3 : 0] = q0.w;
            f1[0] = q1.x; f1[SC > 1 ? 1 : 0] = q1.y; f1[SC > 2 ? 2 : 0] = q1.z; f1[SC > 3 ? 3 : 0] = q1.w;
            f2[0] = q2.x; f2[SC > 1 ? 1 : 0] = q2.y; f2[SC > 2 ? 2 : 0] = q2.z; f2[SC > 3 ? 3 : 0] = q2.w;
            f3[0] = q3.x; f3[SC > 1 ? 1 : 0] = q3.y; f3[SC > 2 ? 2 : 0] = q3.z; f3[SC > 3 ? 3 : 0] = q3.w;
        } else {
#pragma unroll
            for (int c = 0; c < SC; c++) { f0[c] = h0[c]; f1[c] = h1[c]; f2[c] = h2[c]; f3[c] = h3[c]; }
        }
        const float4 fv = __ldg(ybf + by);                              // coefficient * 2^-22, tabulated by the planner
        int v4[4] = {0, 0, 0, 255};
#pragma unroll
        for (int c = 0; c < SC; c++) {
            const float t3 = __fmul_rn(f3[c], fv.w);
            const float t2 = __fadd_rn(__fmul_rn(f2[c], fv.z), t3);
            const float t1 = __fadd_rn(__fmul_rn(f1[c], fv.y), t2);
            const float t0 = __fadd_rn(__fmul_rn(f0[c], fv.x), t1);
            v4[c] = imp_sat8(imp_rint22(t0));
        }
        if (bx * SC + SC > simd_end) {                                  // scalar tail of the row: fixed point (A.4), per byte
            const int b0 = __ldg(yb + by * 4), b1 = __ldg(yb + by * 4 + 1), b2 = __ldg(yb + by * 4 + 2), b3 = __ldg(yb + by * 4 + 3);
#pragma unroll
            for (int c = 0; c < SC; c++)
                if (bx * SC + c >= simd_end)
                    v4[c] = imp_sat8((imp_rint22(f0[c]) * b0 + imp_rint22(f1[c]) * b1 + imp_rint22(f2[c]) * b2 + imp_rint22(f3[c]) * b3 + (1 << 21)) >> 22);
        }
        if (SC == 1) { px[it].b = px[it].g = px[it].r = v4[0]; px[it].a = 255; }
        else { px[it].b = v4[0]; px[it].g = v4[1]; px[it].r = v4[2]; px[it].a = (SC == 4) ? v4[3] : 255; }
    }
    if (nops) imp_run_ops_n<4>(px, oc, bxs, bys, reinterpret_cast<const ImpOp*>(s_ops), nops, s_ops + nops * sizeof(ImpOp), job.wm, job.wm_pitch, job.wm_c);
    if (dc == 4) {
#pragma unroll
        for (int it = 0; it < 4; it++) {
            if (!live[it]) continue;
            const ImpPx& p = px[it];
            uint8_t* d = job.dst + (size_t)(Y0 + (tid >> 5) + it * 8) * job.dst_pitch + (size_t)(X0 + (tid & 31)) * 4;
            *reinterpret_cast<uchar4*>(d) = make_uchar4((unsigned char)p.b, (unsigned char)p.g, (unsigned char)p.r, (unsigned char)p.a);
        }
    } else {                                                            // (the source tile died at the barrier above)
#pragma unroll
        for (int it = 0; it < 4; it++) {
            if (!live[it]) continue;
            const ImpPx& p = px[it];
            uint8_t* d = ostage + ((tid >> 5) + it * 8) * OS + (tid & 31) * 3;
            d[0] = (unsigned char)p.b; d[1] = (unsigned char)p.g; d[2] = (unsigned char)p.r;
        }
        __syncthreads();
        tile_copy_out(ostage, OS, job.dst + (size_t)Y0 * job.dst_pitch + (size_t)X0 * 3, job.dst_pitch, vw * 3, vh, tid, CUBIC_THREADS);
    }
}

}  // namespace imp_tiles
