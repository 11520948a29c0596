// imp_tiles.cuh — shared-memory tile kernels (sm_100a): the source rectangle a CTA needs is staged into
// shared memory by the TMA engine (one cp.async.bulk per source row, completion counted on an mbarrier),
// the gather + op list run out of shared memory / registers, and the result is stored coalesced.
//
// The direct-from-global kernels in imp_kernels.cu remain the general path (any pitch/alignment, any
// footprint); these are selected per job when the source rows are 16-byte addressable
// (imp_tile_eligible) and the footprint fits the shared-memory budget.
#pragma once
#include "imp_gather.cuh"

namespace imp_tiles {

// ---- PTX: mbarrier + 1-D bulk async copy (TMA, SASS UBLKCP) ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, 16-byte aligned on both sides, size a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// u8 -> f32 product without the XU pipe: m = 2^23 + S (PRMT), fma(m, a, -(2^23*a)) == RN(S*a) exactly,
// because m*a - 2^23*a == S*a in exact arithmetic and FMA rounds once.
__device__ __forceinline__ float byte_times(uint32_t word, int sel, float a, float neg_a23) {
    const float m = __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7650 | sel));
    return __fmaf_rn(m, a, neg_a23);
}
__device__ __forceinline__ float u8_times(uint32_t byte, float a, float neg_a23) {
    return __fmaf_rn(__uint_as_float(byte | 0x4B000000u), a, neg_a23);
}

constexpr int TW = 32, TH = 8;          // output tile
constexpr int MAX_SMEM = 96 * 1024;     // source-tile budget per CTA

struct Geom {
    int bx0, by0;            // tile origin in the base frame
    int px0, py0;            // first source pixel column / row staged (window-relative)
    int rows;                // source rows staged
    int row_bytes;           // bytes copied per row (multiple of 16)
    int shift;               // byte offset of pixel px0 inside the first 16-byte block
};

// Stage rows [py0, py0+rows) x pixel columns [px0, px1] of the job's source window. Called by all threads.
template <int SC>
__device__ __forceinline__ void stage_source(const ImpJob& job, const ImpPass* P, Geom& g, int px1, uint8_t* tile, int rs, uint64_t* bar) {
    const uint8_t* w0 = job.src + (size_t)P->sy0 * job.src_pitch + (size_t)P->sx0 * SC;     // window origin
    const uint8_t* first = w0 + (size_t)g.py0 * job.src_pitch + (size_t)g.px0 * SC;
    g.shift = (int)((uintptr_t)first & 15);
    g.row_bytes = (g.shift + (px1 - g.px0 + 1) * SC + 15) & ~15;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid == 0) mbar_expect_tx(bar, (uint32_t)(g.row_bytes * g.rows));
    if (tid < 32) {
        for (int r = tid; r < g.rows; r += 32)
            bulk_g2s(tile + (size_t)r * rs, first - g.shift + (size_t)r * job.src_pitch, (uint32_t)g.row_bytes, bar);
    }
}

// ---- INTER_AREA, fractional scale, from a shared-memory tile ----------------------------
// One thread per output pixel; a warp is one output row of the tile, so the y taps are warp-uniform.
// The x taps of a column are NT contiguous source pixels whose weights sit in registers. Columns with
// fewer taps than NT are zero-padded on the right: the padded products are exactly +0, and adding +0
// leaves a running float sum bit-identical, so the result equals OpenCV's ordered accumulation.
// (A padded tap may read a byte past the staged pixels; any byte converts to a finite float.)
template <int SC, int NT>
__device__ __forceinline__ void area_rows(const uint8_t* __restrict__ col0, int rs, int py0, const ImpAreaTap* __restrict__ yt, int yfirst, int ycount,
                                          const float (&a)[NT], const float (&na)[NT], float (&sum)[SC]) {
    for (int j = 0; j < ycount; j++) {
        const int2 tyr = __ldg(reinterpret_cast<const int2*>(yt + yfirst + j));
        const float beta = __int_as_float(tyr.y);
        const uint8_t* row = col0 + (size_t)(tyr.x - py0) * rs;
        float h[SC];
#pragma unroll
        for (int k = 0; k < NT; k++) {
            if (SC == 4) {
                const uint32_t w = *reinterpret_cast<const uint32_t*>(row + k * 4);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float pr = byte_times(w, c, a[k], na[k]);
                    h[c] = (k == 0) ? pr : __fadd_rn(h[c], pr);
                }
            } else {
#pragma unroll
                for (int c = 0; c < SC; c++) {
                    const float pr = u8_times(row[k * SC + c], a[k], na[k]);
                    h[c] = (k == 0) ? pr : __fadd_rn(h[c], pr);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < SC; c++) {
            const float pr = __fmul_rn(beta, h[c]);
            sum[c] = (j == 0) ? pr : __fadd_rn(sum[c], pr);
        }
    }
}

template <int SC, int NT>
__device__ __forceinline__ void area_pixel(const uint8_t* __restrict__ col0, int rs, int py0, const ImpAreaTap* __restrict__ xt, const ImpRange rx,
                                           const ImpAreaTap* __restrict__ yt, const ImpRange ry, float (&sum)[SC]) {
    float a[NT], na[NT];
#pragma unroll
    for (int k = 0; k < NT; k++) {
        a[k] = (k < rx.count) ? __int_as_float(__ldg(reinterpret_cast<const int2*>(xt + rx.first + k)).y) : 0.0f;
        na[k] = -8388608.0f * a[k];
    }
    area_rows<SC, NT>(col0, rs, py0, yt, ry.first, ry.count, a, na, sum);
}

template <int SC>
__global__ void __launch_bounds__(TW * TH)
imp_area_frac_tile_kernel(const ImpJob* __restrict__ jobs, int first, int count, const ImpJob one) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int jn = blockIdx.y + blockIdx.z * 65535;
    if (jn >= count) return;
    const ImpJob job = jobs ? jobs[first + jn] : one;
    const uint8_t* __restrict__ blob = job.pass;
    const ImpPass* __restrict__ P = reinterpret_cast<const ImpPass*>(blob);
    const int bw = P->bw, bh = P->bh;
    const int tiles_x = (bw + TW - 1) / TW, tiles_y = (bh + TH - 1) / TH;
    if ((int)blockIdx.x >= tiles_x * tiles_y) return;

    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    const int nops = P->nops;
    const int ops_bytes = (nops * (int)sizeof(ImpOp) + P->lut_bytes + 15) & ~15;
    uint8_t* s_ops = smem + 16;
    uint8_t* tile = s_ops + ((ops_bytes + 127) & ~127) + 112;       // keeps the tile 128-byte aligned
    const int tid = threadIdx.y * TW + threadIdx.x;
    if (tid == 0) mbar_init(bar, 1);

    const ImpRange* __restrict__ xr = reinterpret_cast<const ImpRange*>(blob + P->xofs_off);
    const ImpAreaTap* __restrict__ xt = reinterpret_cast<const ImpAreaTap*>(blob + P->xcoef_off);
    const ImpRange* __restrict__ yr = reinterpret_cast<const ImpRange*>(blob + P->yofs_off);
    const ImpAreaTap* __restrict__ yt = reinterpret_cast<const ImpAreaTap*>(blob + P->ycoef_off);
    auto ldr = [](const ImpRange* p) { const int2 v = __ldg(reinterpret_cast<const int2*>(p)); ImpRange r; r.first = v.x; r.count = v.y; return r; };
    auto ldsi = [](const ImpAreaTap* p) { return __ldg(reinterpret_cast<const int*>(p)); };

    Geom g;
    g.bx0 = (blockIdx.x % tiles_x) * TW; g.by0 = (blockIdx.x / tiles_x) * TH;
    const int bx1 = min(g.bx0 + TW, bw) - 1, by1 = min(g.by0 + TH, bh) - 1;
    const ImpRange rx0 = ldr(xr + g.bx0), rx1 = ldr(xr + bx1), ry0 = ldr(yr + g.by0), ry1 = ldr(yr + by1);
    g.px0 = ldsi(xt + rx0.first);
    const int px1 = ldsi(xt + rx1.first + rx1.count - 1);
    g.py0 = ldsi(yt + ry0.first);
    g.rows = ldsi(yt + ry1.first + ry1.count - 1) - g.py0 + 1;
    const int rs = P->tile_rs;
    const int ntx = P->max_xtaps;
    __syncthreads();                                                 // barrier init visible to every thread
    stage_source<SC>(job, P, g, px1, tile, rs, bar);

    // while the TMA engine fills the tile: ops + LUTs into shared memory
    {
        const uint4* gsrc = reinterpret_cast<const uint4*>(blob + P->ops_off);
        uint4* sdst = reinterpret_cast<uint4*>(s_ops);
        for (int i = tid; i < ops_bytes / 16; i += TW * TH) sdst[i] = __ldg(gsrc + i);
    }
    const int bx = min(g.bx0 + (int)threadIdx.x, bw - 1), by = min(g.by0 + (int)threadIdx.y, bh - 1);
    const ImpRange rx = ldr(xr + bx), ry = ldr(yr + by);
    const uint8_t* col0 = tile + g.shift + (ldsi(xt + rx.first) - g.px0) * SC;     // this thread's first tap in tile row 0
    __syncthreads();
    mbar_wait(bar, 0);

    float sum[SC];
    switch (ntx) {                                                   // uniform over the whole pass
        case 1:  area_pixel<SC, 1>(col0, rs, g.py0, xt, rx, yt, ry, sum); break;
        case 2:  area_pixel<SC, 2>(col0, rs, g.py0, xt, rx, yt, ry, sum); break;
        case 3:  area_pixel<SC, 3>(col0, rs, g.py0, xt, rx, yt, ry, sum); break;
        case 4:  area_pixel<SC, 4>(col0, rs, g.py0, xt, rx, yt, ry, sum); break;
        case 5:  area_pixel<SC, 5>(col0, rs, g.py0, xt, rx, yt, ry, sum); break;
        case 6:  area_pixel<SC, 6>(col0, rs, g.py0, xt, rx, yt, ry, sum); break;
        case 7:  area_pixel<SC, 7>(col0, rs, g.py0, xt, rx, yt, ry, sum); break;
        case 8:  area_pixel<SC, 8>(col0, rs, g.py0, xt, rx, yt, ry, sum); break;
        case 9:  area_pixel<SC, 9>(col0, rs, g.py0, xt, rx, yt, ry, sum); break;
        case 10: area_pixel<SC, 10>(col0, rs, g.py0, xt, rx, yt, ry, sum); break;
        case 11: area_pixel<SC, 11>(col0, rs, g.py0, xt, rx, yt, ry, sum); break;
        default: area_pixel<SC, 12>(col0, rs, g.py0, xt, rx, yt, ry, sum); break;
    }
    if ((int)threadIdx.x + g.bx0 >= bw || (int)threadIdx.y + g.by0 >= bh) return;

    ImpPx p;
    if (SC == 1) { p.b = p.g = p.r = imp_sat8(__float2int_rn(sum[0])); p.a = 255; }
    else {
        p.b = imp_sat8(__float2int_rn(sum[0])); p.g = imp_sat8(__float2int_rn(sum[SC > 1 ? 1 : 0])); p.r = imp_sat8(__float2int_rn(sum[SC > 2 ? 2 : 0]));
        p.a = (SC == 4) ? imp_sat8(__float2int_rn(sum[SC - 1])) : 255;
    }
    const int oc = P->oc;
    if (nops) imp_run_ops(p, oc, bx, by, reinterpret_cast<const ImpOp*>(s_ops), nops, s_ops + nops * sizeof(ImpOp), job.wm, job.wm_pitch, job.wm_c);
    int X, Y;
    imp_map_xy(P->out, bx, by, X, Y);
    uint8_t* d = job.dst + (size_t)Y * job.dst_pitch + (size_t)X * oc;
    if (oc == 4) *reinterpret_cast<uchar4*>(d) = make_uchar4((unsigned char)p.b, (unsigned char)p.g, (unsigned char)p.r, (unsigned char)p.a);
    else { d[0] = (unsigned char)p.b; d[1] = (unsigned char)p.g; d[2] = (unsigned char)p.r; }
}

}  // namespace imp_tiles
