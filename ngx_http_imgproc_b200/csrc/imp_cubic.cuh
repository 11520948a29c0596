// imp_cubic.cuh — INTER_CUBIC (cvResize, bridge.c:190-191) from a TMA-staged source tile, sm_100a.
//
// One CTA per T x T output tile (T = 64 or 32, chosen by the planner from the shared-memory footprint), tiled in
// DESTINATION space (a square of the base frame under all eight output orientations, contiguous destination rows):
//   1. ONE TMA box load of the source rectangle the tile's 4x4 footprints touch (clamped to the window: OpenCV
//      replicates the border by clamping tap coordinates, so no fill is needed);
//   2. horizontal pass ONCE per (source row, output column): OpenCV's int32 H = sum p*a (11-bit coefficients), kept as
//      exact floats in shared memory (|H| < 2^22). At 2x that is ~0.55 horizontal rows per output pixel; the per-pixel
//      gather pays 4, the column-run kernel of round 1 paid 0.94;
//   3. per tile row, the four clamped source-row offsets and the four float coefficients are resolved once into shared
//      memory (the first version of this kernel spent a quarter of its instructions clamping them per pixel);
//   4. vertical pass: a thread owns FOUR horizontally adjacent destination pixels per step, in OpenCV's float form (mul
//      and add rounded separately, in its order), rounded and saturated by one cvt.rni.sat.u8.f32; bytes of a row at or
//      beyond simd_end take the integer form of OpenCV's scalar tail;
//   5. op list once over the four pixels, then 4-channel results leave as one 128-bit store per thread, 3-channel results
//      through a shared-memory stage as 16-byte chunks.
#pragma once
#include "imp_blur.cuh"

namespace imp_tiles {

constexpr int CUBIC_THREADS = 256;

// d = c + a.lo16 * b.byte0 + a.hi16 * b.byte1 (lo) / b.byte2, b.byte3 (hi); a signed (coefficients), b unsigned (pixel bytes)
__device__ __forceinline__ int dp2a_lo_su(int a, unsigned b, int c) { int d; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dp2a_hi_su(int a, unsigned b, int c) { int d; asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

__device__ __forceinline__ int rint_sat_u8(float x) {      // sat_u8(rint(x)), round-half-even, one F2I
    unsigned r;
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(x));
    return (int)r;
}

template <int SC, bool LIGHT>
__global__ void __launch_bounds__(CUBIC_THREADS, LIGHT ? 4 : 3)
imp_cubic_tile_kernel(const ImpJob* __restrict__ jobs, int first, int count, const __grid_constant__ ImpJob one) {
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
    const int HRS = IMP_CUBIC_HRS(SC, T);                               // floats per row of the horizontal-pass buffer
    const int bw = P->bw, sw = P->sw, sh = P->sh;
    const ImpFrameMap om = P->out;
    const int tsh = T == 64 ? 6 : 5;                                    // T is 64 or 32
    const int tiles_xd = (om.w + T - 1) >> tsh, tiles_yd = (om.h + T - 1) >> tsh;
    if ((int)blockIdx.x >= tiles_xd * tiles_yd) return;
    const int X0 = ((int)blockIdx.x % tiles_xd) * T, Y0 = ((int)blockIdx.x / tiles_xd) * T;
    const int vw = min(T, om.w - X0), vh = min(T, om.h - Y0);           // valid destination rectangle
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
    uint8_t* tile = s_ops + ((ops_bytes + 127) & ~127);                 // TMA box
    float* hbuf = reinterpret_cast<float*>(tile + ((rs * P->tile_rows + 127) & ~127));      // [tile_rows][HRS]
    int4* rofs = reinterpret_cast<int4*>(hbuf + ((P->tile_rows * HRS + 3) & ~3));   // [T] float offsets of a tile row's four source rows in hbuf
    float4* rcoef = reinterpret_cast<float4*>(rofs + T);                // [T] its four coefficients * 2^-22
    uint8_t* ostage = reinterpret_cast<uint8_t*>(rcoef + T);            // T rows of T*3 bytes (3-channel results)
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
    // the tile rows' lookups, resolved once
    if (tid < th) {
        const int by = y0 + tid;
        const int yo = __ldg(yofs + by) - 1;
        rofs[tid] = make_int4((min(max(yo, 0), sh - 1) - sy_first) * HRS, (min(max(yo + 1, 0), sh - 1) - sy_first) * HRS,
                              (min(max(yo + 2, 0), sh - 1) - sy_first) * HRS, (min(max(yo + 3, 0), sh - 1) - sy_first) * HRS);
        rcoef[tid] = __ldg(reinterpret_cast<const float4*>(blob + P->taps_off) + by);      // coefficient * 2^-22, tabulated by the planner
        IMP_DBG(rofs[tid].x >= 0 && rofs[tid].w < P->tile_rows * HRS && nrows <= P->tile_rows, 3);
    }
    __syncthreads();                                                    // the mbarrier is initialised, ops and row lookups are staged
    mbar_wait(bar, 0);
    // ---- horizontal pass: a thread owns one output column of the tile (two when T = 64) and walks the source rows ----
    for (int lx = tid & 31; lx < tw; lx += 32) {
        const int bx = x0 + lx;
        const int xo = __ldg(xofs + bx) - 1;
        int so[4];                                                      // byte offsets of the four taps inside a tile row
#pragma unroll
        for (int t = 0; t < 4; t++) so[t] = col_off + (min(max(xo + t, 0), sw - 1) - sx_first) * SC;
        const int2 av = __ldg(reinterpret_cast<const int2*>(xa + bx * 4));   // 4 shorts, 8-byte aligned
        const int a0 = (short)(av.x & 0xffff), a1 = av.x >> 16, a2 = (short)(av.y & 0xffff), a3 = av.y >> 16;
        IMP_DBG(so[0] >= 0 && so[3] + SC <= rs, 2);
        for (int r = tid >> 5; r < nrows; r += CUBIC_THREADS / 32) {
            const uint8_t* row = tile + r * rs;
            int hsum[SC];
            if (SC == 4) {
                const uint32_t p0 = *reinterpret_cast<const uint32_t*>(row + so[0]), p1 = *reinterpret_cast<const uint32_t*>(row + so[1]);
                const uint32_t p2 = *reinterpret_cast<const uint32_t*>(row + so[2]), p3 = *reinterpret_cast<const uint32_t*>(row + so[3]);
                // byte-transpose pixel pairs (4 PRMT), then each channel is two mixed-sign dp2a: 12 instructions for the 16 MACs
                const unsigned t01 = __byte_perm(p0, p1, 0x5140), u01 = __byte_perm(p0, p1, 0x7362);     // {c0:p0,p1 | c1:p0,p1}, {c2 | c3}
                const unsigned t23 = __byte_perm(p2, p3, 0x5140), u23 = __byte_perm(p2, p3, 0x7362);
                hsum[0] = dp2a_lo_su(av.x, t01, dp2a_lo_su(av.y, t23, 0));
                hsum[SC > 1 ? 1 : 0] = dp2a_hi_su(av.x, t01, dp2a_hi_su(av.y, t23, 0));
                hsum[SC > 2 ? 2 : 0] = dp2a_lo_su(av.x, u01, dp2a_lo_su(av.y, u23, 0));
                hsum[SC > 3 ? 3 : 0] = dp2a_hi_su(av.x, u01, dp2a_hi_su(av.y, u23, 0));
            } else {
#pragma unroll
                for (int c = 0; c < SC; c++)
                    hsum[c] = (int)row[so[0] + c] * a0 + (int)row[so[1] + c] * a1 + (int)row[so[2] + c] * a2 + (int)row[so[3] + c] * a3;
            }
            float* hp = hbuf + r * HRS + IMP_CUBIC_POS(lx, T) * SC;
            if (SC == 4) *reinterpret_cast<float4*>(hp) = make_float4(imp_i2f22(hsum[0]), imp_i2f22(hsum[SC > 1 ? 1 : 0]), imp_i2f22(hsum[SC > 2 ? 2 : 0]), imp_i2f22(hsum[SC > 3 ? 3 : 0]));
            else {
#pragma unroll
                for (int c = 0; c < SC; c++) hp[c] = imp_i2f22(hsum[c]);
            }
        }
    }
    __syncthreads();

    // ---- vertical pass + op list: a thread owns 4 horizontally adjacent destination pixels per step ----
    const short* __restrict__ yb = reinterpret_cast<const short*>(blob + P->ycoef_off);
    const int oc = P->oc, dc = P->dc, simd_end = P->simd_end;
    const int OS = T * 3;                                               // out-stage row stride (3-channel results)
    const int qsh = tsh - 2;                                            // log2 of the 4-pixel groups per tile row
    const bool vec16 = dc == 4 && ((reinterpret_cast<uintptr_t>(job.dst) | (unsigned)job.dst_pitch) & 15) == 0;
    const int dstep = om.flipx ? -1 : 1, dbx = om.swap ? 0 : dstep, dby = om.swap ? dstep : 0;
    for (int item = tid; item < (T << qsh); item += CUBIC_THREADS) {
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
            const int4 ro = rofs[ly];
            const float4 fv = rcoef[ly];
            const float* hb = hbuf + IMP_CUBIC_POS(lx, T) * SC;
            float f0[SC], f1[SC], f2[SC], f3[SC];
            if (SC == 4) {
                const float4 q0 = *reinterpret_cast<const float4*>(hb + ro.x), q1 = *reinterpret_cast<const float4*>(hb + ro.y);
                const float4 q2 = *reinterpret_cast<const float4*>(hb + ro.z), q3 = *reinterpret_cast<const float4*>(hb + ro.w);
                f0[0] = q0.x; f0[SC > 1 ? 1 : 0] = q0.y; f0[SC > 2 ? 2 : 0] = q0.z; f0[SC > 3 ? 3 : 0] = q0.w;
                f1[0] = q1.x; f1[SC > 1 ? 1 : 0] = q1.y; f1[SC > 2 ? 2 : 0] = q1.z; f1[SC > 3 ? 3 : 0] = q1.w;
                f2[0] = q2.x; f2[SC > 1 ? 1 : 0] = q2.y; f2[SC > 2 ? 2 : 0] = q2.z; f2[SC > 3 ? 3 : 0] = q2.w;
                f3[0] = q3.x; f3[SC > 1 ? 1 : 0] = q3.y; f3[SC > 2 ? 2 : 0] = q3.z; f3[SC > 3 ? 3 : 0] = q3.w;
            } else {
#pragma unroll
                for (int c = 0; c < SC; c++) { f0[c] = hb[ro.x + c]; f1[c] = hb[ro.y + c]; f2[c] = hb[ro.z + c]; f3[c] = hb[ro.w + c]; }
            }
            int v4[4] = {0, 0, 0, 255};
#pragma unroll
            for (int c = 0; c < SC; c++) {
                const float t3 = __fmul_rn(f3[c], fv.w);
                const float t2 = __fadd_rn(__fmul_rn(f2[c], fv.z), t3);
                const float t1 = __fadd_rn(__fmul_rn(f1[c], fv.y), t2);
                const float t0 = __fadd_rn(__fmul_rn(f0[c], fv.x), t1);
                v4[c] = rint_sat_u8(t0);
            }
            if (bx * SC + SC > simd_end) {                              // scalar tail of the row: fixed point (A.4), per byte
                const int b0 = __ldg(yb + by * 4), b1 = __ldg(yb + by * 4 + 1), b2 = __ldg(yb + by * 4 + 2), b3 = __ldg(yb + by * 4 + 3);
#pragma unroll
                for (int c = 0; c < SC; c++)
                    if (bx * SC + c >= simd_end)
                        v4[c] = imp_sat8((imp_rint22(f0[c]) * b0 + imp_rint22(f1[c]) * b1 + imp_rint22(f2[c]) * b2 + imp_rint22(f3[c]) * b3 + (1 << 21)) >> 22);
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
            uint32_t* d = reinterpret_cast<uint32_t*>(ostage + Yl * OS + Xq * 3);          // 12 bytes = three aligned words
            d[0] = (w[0] & 0xFFFFFFu) | (w[1] << 24);
            d[1] = ((w[1] >> 8) & 0xFFFFu) | (w[2] << 16);
            d[2] = ((w[2] >> 16) & 0xFFu) | (w[3] << 8);
        }
    }
    if (dc != 4) {
        __syncthreads();
        tile_copy_out(ostage, OS, job.dst + (size_t)Y0 * job.dst_pitch + (size_t)X0 * 3, job.dst_pitch, vw * 3, vh, tid, CUBIC_THREADS);
    }
}

}  // namespace imp_tiles
