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

// ---- INTER_AREA, fractional scale, from a shared-memory tile ----------------------------
// One thread per output pixel; a warp is one output row of the tile, so the y taps are warp-uniform.
// area_rows() accumulates one output pixel from the staged rows.
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

// Persistent strip kernel. A CTA owns a strip of 32 output columns of one job and walks down its 8-row
// tiles. Warp 8 is the TMA producer: it stages the source rows of tile t+1 into the other half of a
// two-stage shared-memory ring while warps 0-7 (one output row each) consume tile t. full[]/empty[]
// mbarriers carry the hand-off; the per-thread x weights are loaded once per strip.
//   blob tables (host, imp_planner.cpp): xtile[tx] = {first source px, last source px} of tile column tx,
//   ytile[ty] = {first source row, number of source rows} of tile row ty.
constexpr int STRIP_CONSUMERS = TW * TH;          // 256 threads, 8 warps
constexpr int STRIP_THREADS = STRIP_CONSUMERS + 32;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int SC, int NT, int NSTAGE>
__device__ __forceinline__ void area_strip_consume(const ImpJob& job, const ImpPass* __restrict__ P, const uint8_t* __restrict__ blob,
                                                   const uint8_t* tile0, int stage_bytes, uint64_t* full, uint64_t* empty,
                                                   const uint8_t* s_ops, int nops, int bx0, int col_off, int tiles_y) {
    const ImpRange* __restrict__ xr = reinterpret_cast<const ImpRange*>(blob + P->xofs_off);
    const ImpAreaTap* __restrict__ xt = reinterpret_cast<const ImpAreaTap*>(blob + P->xcoef_off);
    const ImpRange* __restrict__ yr = reinterpret_cast<const ImpRange*>(blob + P->yofs_off);
    const ImpAreaTap* __restrict__ yt = reinterpret_cast<const ImpAreaTap*>(blob + P->ycoef_off);
    const int2* __restrict__ ytile = reinterpret_cast<const int2*>(blob + P->ytile_off);
    const int bw = P->bw, bh = P->bh, rs = P->tile_rs, oc = P->oc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool in_x = bx0 + lane < bw;
    const int bx = min(bx0 + lane, bw - 1);
    const int2 rxv = __ldg(reinterpret_cast<const int2*>(xr + bx));
    float a[NT], na[NT];
#pragma unroll
    for (int k = 0; k < NT; k++) {
        a[k] = (k < rxv.y) ? __int_as_float(__ldg(reinterpret_cast<const int2*>(xt + rxv.x + k)).y) : 0.0f;
        na[k] = -8388608.0f * a[k];
    }
    const int my_off = col_off + __ldg(reinterpret_cast<const int*>(xt + rxv.x)) * SC;    // byte offset of my first tap in a tile row
    const ImpFrameMap om = P->out;

    int stage = 0, phase = 0;
    for (int t = 0; t < tiles_y; t++) {
        const int by = min(t * TH + warp, bh - 1);
        const bool in_y = t * TH + warp < bh;
        const int2 ryv = __ldg(reinterpret_cast<const int2*>(yr + by));
        const int py0 = __ldg(ytile + t).x;
        mbar_wait(full + stage, phase);
        float sum[SC];
        area_rows<SC, NT>(tile0 + stage * stage_bytes + my_off, rs, py0, yt, ryv.x, ryv.y, a, na, sum);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + stage);                     // this warp is done with the stage
        if (in_x && in_y) {
            ImpPx p;
            if (SC == 1) { p.b = p.g = p.r = imp_sat8(__float2int_rn(sum[0])); p.a = 255; }
            else {
                p.b = imp_sat8(__float2int_rn(sum[0])); p.g = imp_sat8(__float2int_rn(sum[SC > 1 ? 1 : 0])); p.r = imp_sat8(__float2int_rn(sum[SC > 2 ? 2 : 0]));
                p.a = (SC == 4) ? imp_sat8(__float2int_rn(sum[SC - 1])) : 255;
            }
            if (nops) imp_run_ops(p, oc, bx, by, reinterpret_cast<const ImpOp*>(s_ops), nops, s_ops + nops * sizeof(ImpOp), job.wm, job.wm_pitch, job.wm_c);
            int X, Y;
            imp_map_xy(om, bx, by, X, Y);
            uint8_t* d = job.dst + (size_t)Y * job.dst_pitch + (size_t)X * oc;
            if (oc == 4) *reinterpret_cast<uchar4*>(d) = make_uchar4((unsigned char)p.b, (unsigned char)p.g, (unsigned char)p.r, (unsigned char)p.a);
            else { d[0] = (unsigned char)p.b; d[1] = (unsigned char)p.g; d[2] = (unsigned char)p.r; }
        }
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
    }
}

template <int SC, int NSTAGE>
__global__ void __launch_bounds__(STRIP_THREADS, 3)
imp_area_frac_strip_kernel(const ImpJob* __restrict__ jobs, int first, int count, const ImpJob one) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int jn = blockIdx.y + blockIdx.z * 65535;
    if (jn >= count) return;
    const ImpJob job = jobs ? jobs[first + jn] : one;
    const uint8_t* __restrict__ blob = job.pass;
    const ImpPass* __restrict__ P = reinterpret_cast<const ImpPass*>(blob);
    const int bw = P->bw, bh = P->bh;
    const int tiles_x = (bw + TW - 1) / TW, tiles_y = (bh + TH - 1) / TH;
    if ((int)blockIdx.x >= tiles_x) return;

    uint64_t* full = reinterpret_cast<uint64_t*>(smem);               // [NSTAGE]
    uint64_t* empty = full + NSTAGE;                                  // [NSTAGE]
    const int nops = P->nops;
    const int ops_bytes = (nops * (int)sizeof(ImpOp) + P->lut_bytes + 15) & ~15;
    uint8_t* s_ops = smem + 64;
    uint8_t* tile0 = s_ops + ((ops_bytes + 127) & ~127) + 64;         // 128-byte aligned
    const int rs = P->tile_rs;
    const int stage_bytes = (rs * P->tile_rows + 64 + 127) & ~127;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; i++) { mbar_init(full + i, 1); mbar_init(empty + i, TH); }
    }
    {
        const uint4* gsrc = reinterpret_cast<const uint4*>(blob + P->ops_off);
        uint4* sdst = reinterpret_cast<uint4*>(s_ops);
        for (int i = tid; i < ops_bytes / 16; i += STRIP_THREADS) sdst[i] = __ldg(gsrc + i);
    }
    const int2 xt_tile = __ldg(reinterpret_cast<const int2*>(blob + P->xtile_off) + blockIdx.x);     // {px0, px1}
    const uint8_t* w0 = job.src + (size_t)P->sy0 * job.src_pitch + (size_t)P->sx0 * SC;            // window origin
    const uint8_t* col_first = w0 + (size_t)xt_tile.x * SC;
    const int shift = (int)((uintptr_t)col_first & 15);
    const int row_bytes = (shift + (xt_tile.y - xt_tile.x + 1) * SC + 15) & ~15;
    __syncthreads();

    if (tid >= STRIP_CONSUMERS) {
        // ---- TMA producer warp ----
        const int lane = tid & 31;
        const int2* __restrict__ ytile = reinterpret_cast<const int2*>(blob + P->ytile_off);
        int stage = 0, phase = 0;
        for (int t = 0; t < tiles_y; t++) {
            const int2 yt_tile = __ldg(ytile + t);                                                  // {py0, rows}
            mbar_wait(empty + stage, phase ^ 1);
            if (lane == 0) mbar_expect_tx(full + stage, (uint32_t)(row_bytes * yt_tile.y));
            __syncwarp();
            uint8_t* dst = tile0 + stage * stage_bytes;
            const uint8_t* src = col_first - shift + (size_t)yt_tile.x * job.src_pitch;
            for (int r = lane; r < yt_tile.y; r += 32)
                bulk_g2s(dst + (size_t)r * rs, src + (size_t)r * job.src_pitch, (uint32_t)row_bytes, full + stage);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        return;
    }
    // ---- consumers ----
    const int col_off = shift - xt_tile.x * SC;                       // tile byte offset of source pixel 0
    const int bx0 = blockIdx.x * TW;
    switch (P->max_xtaps) {                                           // uniform over the pass
        case 1:  area_strip_consume<SC, 1, NSTAGE>(job, P, blob, tile0, stage_bytes, full, empty, s_ops, nops, bx0, col_off, tiles_y); break;
        case 2:  area_strip_consume<SC, 2, NSTAGE>(job, P, blob, tile0, stage_bytes, full, empty, s_ops, nops, bx0, col_off, tiles_y); break;
        case 3:  area_strip_consume<SC, 3, NSTAGE>(job, P, blob, tile0, stage_bytes, full, empty, s_ops, nops, bx0, col_off, tiles_y); break;
        case 4:  area_strip_consume<SC, 4, NSTAGE>(job, P, blob, tile0, stage_bytes, full, empty, s_ops, nops, bx0, col_off, tiles_y); break;
        case 5:  area_strip_consume<SC, 5, NSTAGE>(job, P, blob, tile0, stage_bytes, full, empty, s_ops, nops, bx0, col_off, tiles_y); break;
        case 6:  area_strip_consume<SC, 6, NSTAGE>(job, P, blob, tile0, stage_bytes, full, empty, s_ops, nops, bx0, col_off, tiles_y); break;
        case 7:  area_strip_consume<SC, 7, NSTAGE>(job, P, blob, tile0, stage_bytes, full, empty, s_ops, nops, bx0, col_off, tiles_y); break;
        case 8:  area_strip_consume<SC, 8, NSTAGE>(job, P, blob, tile0, stage_bytes, full, empty, s_ops, nops, bx0, col_off, tiles_y); break;
        case 9:  area_strip_consume<SC, 9, NSTAGE>(job, P, blob, tile0, stage_bytes, full, empty, s_ops, nops, bx0, col_off, tiles_y); break;
        case 10: area_strip_consume<SC, 10, NSTAGE>(job, P, blob, tile0, stage_bytes, full, empty, s_ops, nops, bx0, col_off, tiles_y); break;
        case 11: area_strip_consume<SC, 11, NSTAGE>(job, P, blob, tile0, stage_bytes, full, empty, s_ops, nops, bx0, col_off, tiles_y); break;
        default: area_strip_consume<SC, 12, NSTAGE>(job, P, blob, tile0, stage_bytes, full, empty, s_ops, nops, bx0, col_off, tiles_y); break;
    }
}

}  // namespace imp_tiles
