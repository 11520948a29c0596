// imp_gather.cuh — the per-output-pixel gathers (crop/NN/area/cubic/linear), written against an
// abstract source accessor so the same arithmetic serves the direct-from-global kernels and the
// shared-memory tile kernels, and can be instantiated on the host for the CPU unit test.
//
// Arithmetic follows SURVEY Appendix A (OpenCV's algorithms as the reference's cvResize call,
// bridge.c:191, executes them); tables are built on the host by imp_planner.cpp with the same double
// maths OpenCV uses.
#pragma once
#include "imp_pixel.cuh"

struct ImpAreaTap { int si; float a; };      // source index (pixels, window-relative), weight
struct ImpRange   { int first, count; };     // taps of one output column/row

// Accessor over global memory: window-relative pixel coordinates.
template <int SC>
struct ImpSrcGlobal {
    const uint8_t* base;   // already offset to the window origin
    int pitch;
    IMP_HD int at(int x, int y, int c) const { return base[(size_t)y * pitch + x * SC + c]; }
    IMP_HD void px(int x, int y, int* v) const {
        const uint8_t* p = base + (size_t)y * pitch + x * SC;
#if defined(__CUDA_ARCH__)
        if (SC == 4) { uchar4 q = *reinterpret_cast<const uchar4*>(p); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; return; }
#endif
#pragma unroll
        for (int c = 0; c < SC; c++) v[c] = p[c];
    }
};

template <int SC, class Src>
IMP_HD void imp_gather_copy(const Src& S, int bx, int by, int* v) { S.px(bx, by, v); }

template <int SC, class Src>
IMP_HD void imp_gather_nn(const Src& S, const int* xofs, const int* yofs, int bx, int by, int* v) {
    S.px(xofs[bx], yofs[by], v);
}

// A.2: integer box; 2x2 -> (s+2)>>2, else rint(f32(s) * f32(1/(nx*ny))).
template <int SC, class Src>
IMP_HD void imp_gather_area_int(const Src& S, int nx, int ny, float scale, int bx, int by, int* v) {
    int sum[SC];
#pragma unroll
    for (int c = 0; c < SC; c++) sum[c] = 0;
    int x0 = bx * nx, y0 = by * ny;
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            int t[SC];
            S.px(x0 + i, y0 + j, t);
#pragma unroll
            for (int c = 0; c < SC; c++) sum[c] += t[c];
        }
    if (nx == 2 && ny == 2) {
#pragma unroll
        for (int c = 0; c < SC; c++) v[c] = (sum[c] + 2) >> 2;
    } else {
#pragma unroll
        for (int c = 0; c < SC; c++) v[c] = imp_sat8(IMP_RINT(IMP_FMUL((float)sum[c], scale)));
    }
}

// A.3: ordered float32 accumulation, x taps inside y taps, no FMA.
template <int SC, class Src>
IMP_HD void imp_gather_area_frac(const Src& S, const ImpRange* xr, const ImpAreaTap* xt,
                                 const ImpRange* yr, const ImpAreaTap* yt, int bx, int by, int* v) {
    ImpRange rx = xr[bx], ry = yr[by];
    float sum[SC];
    for (int j = 0; j < ry.count; j++) {
        ImpAreaTap ty = yt[ry.first + j];
        float buf[SC];
        for (int k = 0; k < rx.count; k++) {
            ImpAreaTap tx = xt[rx.first + k];
            int t[SC];
            S.px(tx.si, ty.si, t);
#pragma unroll
            for (int c = 0; c < SC; c++) {
                float pr = IMP_FMUL((float)t[c], tx.a);
                buf[c] = (k == 0) ? pr : IMP_FADD(buf[c], pr);
            }
        }
#pragma unroll
        for (int c = 0; c < SC; c++) {
            float pr = IMP_FMUL(ty.a, buf[c]);
            sum[c] = (j == 0) ? pr : IMP_FADD(sum[c], pr);
        }
    }
#pragma unroll
    for (int c = 0; c < SC; c++) v[c] = imp_sat8(IMP_RINT(sum[c]));
}

// A.4 cubic: int32 horizontal pass with 11-bit coefficients, vertical pass in OpenCV's float SIMD
// form for row bytes < simd_end and in fixed point for the scalar tail.
template <int SC, class Src>
IMP_HD void imp_gather_cubic(const Src& S, int sw, int sh, const int* xofs, const short* xa, const int* yofs,
                             const short* yb, int simd_end, int bx, int by, int* v) {
    int sx[4], sy[4];
    int x0 = xofs[bx] - 1, y0 = yofs[by] - 1;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        sx[t] = imp_min(imp_max(x0 + t, 0), sw - 1);
        sy[t] = imp_min(imp_max(y0 + t, 0), sh - 1);
    }
    int a0 = xa[bx * 4], a1 = xa[bx * 4 + 1], a2 = xa[bx * 4 + 2], a3 = xa[bx * 4 + 3];
    int H[4][SC];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int p0[SC], p1[SC], p2[SC], p3[SC];
        S.px(sx[0], sy[k], p0); S.px(sx[1], sy[k], p1); S.px(sx[2], sy[k], p2); S.px(sx[3], sy[k], p3);
#pragma unroll
        for (int c = 0; c < SC; c++) H[k][c] = p0[c] * a0 + p1[c] * a1 + p2[c] * a2 + p3[c] * a3;
    }
    int b0 = yb[by * 4], b1 = yb[by * 4 + 1], b2 = yb[by * 4 + 2], b3 = yb[by * 4 + 3];
    const float sc = 1.0f / 4194304.0f;               // 2^-22, exact
    float f0 = IMP_FMUL(imp_i2f22(b0), sc), f1 = IMP_FMUL(imp_i2f22(b1), sc), f2 = IMP_FMUL(imp_i2f22(b2), sc), f3 = IMP_FMUL(imp_i2f22(b3), sc);
#pragma unroll
    for (int c = 0; c < SC; c++) {
        if (bx * SC + c < simd_end) {
            // |H| <= 255 * 2048 * 1.3 < 2^22 and |t0| < 2^22: the XU-free conversions are exact
            float t3 = IMP_FMUL(imp_i2f22(H[3][c]), f3);
            float t2 = IMP_FADD(IMP_FMUL(imp_i2f22(H[2][c]), f2), t3);
            float t1 = IMP_FADD(IMP_FMUL(imp_i2f22(H[1][c]), f1), t2);
            float t0 = IMP_FADD(IMP_FMUL(imp_i2f22(H[0][c]), f0), t1);
            v[c] = imp_sat8(imp_rint22(t0));
        } else {
            v[c] = imp_sat8((H[0][c] * b0 + H[1][c] * b1 + H[2][c] * b2 + H[3][c] * b3 + (1 << 21)) >> 22);
        }
    }
}

// A.4 linear: (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2.
template <int SC, class Src>
IMP_HD void imp_gather_linear(const Src& S, int sw, int sh, const int* xofs, const short* xa, const int* yofs,
                              const short* yb, int bx, int by, int* v) {
    int x0 = imp_min(imp_max(xofs[bx], 0), sw - 1), x1 = imp_min(imp_max(xofs[bx] + 1, 0), sw - 1);
    int y0 = imp_min(imp_max(yofs[by], 0), sh - 1), y1 = imp_min(imp_max(yofs[by] + 1, 0), sh - 1);
    int a0 = xa[bx * 2], a1 = xa[bx * 2 + 1], b0 = yb[by * 2], b1 = yb[by * 2 + 1];
    int p00[SC], p01[SC], p10[SC], p11[SC];
    S.px(x0, y0, p00); S.px(x1, y0, p01); S.px(x0, y1, p10); S.px(x1, y1, p11);
#pragma unroll
    for (int c = 0; c < SC; c++) {
        int h0 = p00[c] * a0 + p01[c] * a1, h1 = p10[c] * a0 + p11[c] * a1;
        v[c] = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2 & 255;
    }
}
