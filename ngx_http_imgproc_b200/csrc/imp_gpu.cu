// imp_gpu.cu — the C ABI of libimp_gpu.so (include/imp_gpu.h): device contexts, plan upload,
// batches, the end-to-end host paths (pinned staging + streams) and the multi-GPU farm.
// No CPU fallback exists anywhere in this file: if CUDA is unusable every compute call fails with
// IMP_ERROR_GPU and a message in imp_gpu_last_error().
#include "imp_internal.h"
#include <cuda.h>
#include <stdio.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <new>
#include <thread>

namespace {

constexpr int MAX_DEV = 16;
thread_local char t_err[512] = "";
thread_local int t_dev = -1;

struct Lane {            // one in-flight host job of the end-to-end path
    cudaStream_t st = nullptr;
    uint8_t* d_in = nullptr; size_t in_cap = 0; uint8_t* d_out = nullptr; size_t out_cap = 0;
    uint8_t* h_in = nullptr; size_t hin_cap = 0; uint8_t* h_out = nullptr; size_t hout_cap = 0;
    uint8_t* d_scratch = nullptr; size_t scratch_cap = 0;
    uint8_t* d_stage = nullptr; size_t stage_cap = 0;          // linear landing zone of the host rows (see issue())
    int pending = -1; bool out_staged = false;
};
struct DevCtx {
    bool ready = false;
    cudaStream_t stream = nullptr;
    std::vector<Lane> lanes;       // persistent staging, reused across calls
    std::mutex run_mu;             // serialises users of `lanes`
};
DevCtx g_dev[MAX_DEV];
std::mutex g_mu;

int fail(cudaError_t e, const char* what, int line) {
    snprintf(t_err, sizeof t_err, "%s failed at imp_gpu.cu:%d: %s", what, line, cudaGetErrorString(e));
    return IMP_ERROR_GPU;
}
int fail_msg(const char* msg) { snprintf(t_err, sizeof t_err, "%s", msg); return IMP_ERROR_GPU; }

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(e__, #call, __LINE__); } while (0)

int cur_dev() {
    if (t_dev >= 0) return t_dev;
    std::lock_guard<std::mutex> lk(g_mu);
    for (int d = 0; d < MAX_DEV; d++) if (g_dev[d].ready) { t_dev = d; return d; }
    return -1;
}

int bind() {          // make the thread's device current
    int d = cur_dev();
    if (d < 0) return fail_msg("imp_gpu_init() has not been called (no CUDA device initialised; there is no CPU fallback)");
    CK(cudaSetDevice(d));
    return IMP_OK;
}

cudaStream_t pick_stream(void* s) { return s ? (cudaStream_t)s : g_dev[t_dev].stream; }

int align16(int v) { return (v + 15) & ~15; }

// Uploads the plan's pass blobs and watermark to the current device (once).
int plan_to_device(imp_gpu_plan* plan) {
    const int d = t_dev;
    std::lock_guard<std::mutex> lk(g_mu);
    imp_gpu_plan::Dev& pd = plan->dev[d];
    if (pd.ready) return IMP_OK;
    pd.pass_blobs.assign(plan->passes.size(), nullptr);
    for (size_t i = 0; i < plan->passes.size(); i++) {
        std::vector<uint8_t> b = plan->passes[i].blob;            // per-device copy: vignette table pointers are patched in
        const ImpPass& h = plan->passes[i].hdr;
        ImpOp* ops = reinterpret_cast<ImpOp*>(b.data() + h.ops_off);
        for (int k = 0; k < h.nops; k++) {
            if (ops[k].kind != IMP_OP_VIGNETTE || ops[k].i[2] <= 0) continue;
            float* tab = nullptr;
            CK(cudaMalloc((void**)&tab, (size_t)ops[k].i[2] * sizeof(float)));
            pd.vignette_tabs.push_back(tab);
            CK(imp_build_vignette_table(tab, ops[k].i[2], ops[k].f[0], ops[k].f[1], g_dev[d].stream));
            const unsigned long long v = (unsigned long long)(uintptr_t)tab;
            ops[k].i[4] = (int)(unsigned)(v & 0xffffffffu); ops[k].i[5] = (int)(unsigned)(v >> 32);
        }
        CK(cudaMalloc((void**)&pd.pass_blobs[i], b.size()));
        CK(cudaMemcpy(pd.pass_blobs[i], b.data(), b.size(), cudaMemcpyHostToDevice));
    }
    CK(cudaStreamSynchronize(g_dev[d].stream));
    if (!plan->wm_pixels.empty()) {
        pd.wm_pitch = align16(plan->wm_w * plan->wm_c);
        CK(cudaMalloc((void**)&pd.wm, (size_t)pd.wm_pitch * plan->wm_h));
        CK(cudaMemcpy2D(pd.wm, pd.wm_pitch, plan->wm_pixels.data(), (size_t)plan->wm_w * plan->wm_c,
                        (size_t)plan->wm_w * plan->wm_c, plan->wm_h, cudaMemcpyHostToDevice));
    }
    pd.ready = true;
    return IMP_OK;
}

void plan_free_device(imp_gpu_plan* plan) {
    for (int d = 0; d < MAX_DEV; d++) {
        imp_gpu_plan::Dev& pd = plan->dev[d];
        if (!pd.ready) continue;
        if (cudaSetDevice(d) != cudaSuccess) continue;
        for (uint8_t* p : pd.pass_blobs) cudaFree(p);
        if (pd.wm) cudaFree(pd.wm);
        for (float* t : pd.vignette_tabs) cudaFree(t);
        pd.vignette_tabs.clear(); pd.wm = nullptr;
        pd.ready = false;
    }
}

int ops_smem(const ImpPass& h) { return std::max(16, h.nops * (int)sizeof(ImpOp) + h.lut_bytes); }

}  // namespace

// ---- batch ---------------------------------------------------------------------------------------------
struct imp_gpu_batch {
    struct Item { imp_gpu_plan* plan; const uint8_t* src; int sp; uint8_t* dst; int dp; };
    std::vector<Item> items;
    int dev = -1;
    bool dirty = true;
    // compiled form
    std::vector<ImpJob> h_jobs;
    ImpJob* d_jobs = nullptr; size_t jobs_cap = 0;
    uint8_t* d_scratch = nullptr; size_t scratch_cap = 0;
    struct Step { bool generic_blur; ImpLaunchGroup g; int job; ImpPass hdr; size_t scratch_off; int smem; };
    std::vector<Step> steps;
    unsigned long long algo_bytes = 0;
    int launches = 0;
};

namespace {

int pass_pitch(const ImpHostPass& hp) { return align16(hp.out_w * hp.out_c); }

// Scratch a plan needs on the device: intermediates between passes + the u16 plane of each generic blur.
// off[k] = offset of pass k's output (non-final passes); blur_off[k] = offset of pass k's u16 plane.
int pick_variant(const ImpPass& h, const ImpJob& j);
int variant_param(const ImpPass& h, int variant);
int variant_smem(const ImpPass& h, int variant, int param);
constexpr int kTileSmemLimit = 200 * 1024;     // cudaFuncAttributeMaxDynamicSharedMemorySize of every tile kernel (imp_kernels.cu)

// `src`/`sp`: the job's source (null = one of the library's own aligned buffers). The u16 plane is only reserved for a
// blur that cannot take the fused tile kernel (radius > 12, or a pass-0 source whose rows are not 16-byte addressable).
size_t plan_scratch_layout(const imp_gpu_plan* p, size_t base, std::vector<size_t>& off, std::vector<size_t>& blur_off,
                           const uint8_t* src = nullptr, int sp = 0) {
    size_t cur = base;
    auto take = [&](size_t bytes) { size_t o = cur; cur += (bytes + 255) & ~size_t(255); return o; };
    off.assign(p->passes.size(), 0); blur_off.assign(p->passes.size(), 0);
    for (size_t k = 0; k < p->passes.size(); k++) {
        const ImpHostPass& hp = p->passes[k];
        if (k + 1 < p->passes.size()) off[k] = take((size_t)pass_pitch(hp) * hp.out_h);
        if (hp.hdr.kind == IMP_G_BLUR) {
            ImpJob probe{};
            probe.src = (k == 0 && src) ? src : reinterpret_cast<const uint8_t*>(uintptr_t(256));
            probe.src_pitch = (k == 0 && src) ? sp : 16;
            if (pick_variant(hp.hdr, probe) == 0) blur_off[k] = take((size_t)hp.hdr.sw * hp.hdr.sh * hp.hdr.sc * 2);
        }
    }
    return cur;
}

ImpJob make_job(const imp_gpu_plan* p, int d, size_t k, const uint8_t* src, int sp, uint8_t* dst, int dp,
                uint8_t* scratch, const std::vector<size_t>& off) {
    const imp_gpu_plan::Dev& pd = p->dev[d];
    ImpJob j{};
    j.pass = pd.pass_blobs[k];
    j.wm = pd.wm; j.wm_pitch = pd.wm_pitch; j.wm_c = p->wm_c;
    if (k == 0) { j.src = src; j.src_pitch = sp; }
    else { j.src = scratch + off[k - 1]; j.src_pitch = pass_pitch(p->passes[k - 1]); }
    if (k + 1 == p->passes.size()) { j.dst = dst; j.dst_pitch = dp; }
    else { j.dst = scratch + off[k]; j.dst_pitch = pass_pitch(p->passes[k]); }
    return j;
}

int pass_tiles(const ImpPass& h) { return ((h.bw + 31) / 32) * ((h.bh + 31) / 32); }      // imp_pass_kernel: 32 x (8 x PASS_PPT) pixels per CTA

// Tile (shared-memory, TMA-staged) variant: needs 16-byte addressable source rows — pitch % 16 == 0 and
// either the image rows or the window rows start 16-byte aligned (imp_tiles.cuh) — and a footprint that fits.
// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// Tensor map over the job's source window for the strip kernels: 8-byte elements, rows = window rows, origin =
// the window's first byte aligned down to 16; box = tile_rs bytes x tile_rows rows. Out-of-range box parts are
// zero-filled by the TMA unit, so edge tiles never touch memory outside the window's rows.
int encode_job_tmap(const ImpPass& h, ImpJob& j) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return fail_msg("cuTensorMapEncodeTiled is unavailable in this driver");
    const uintptr_t win = (uintptr_t)j.src + (size_t)h.sy0 * j.src_pitch + (size_t)h.sx0 * h.sc;
    const uintptr_t a0 = win & ~uintptr_t(15);
    j.tm_x0 = (int)(win - a0);
    const cuuint64_t gdim[2] = {(cuuint64_t)((j.tm_x0 + (size_t)h.sw * h.sc + 7) / 8), (cuuint64_t)h.sh};
    const cuuint64_t gstride[1] = {(cuuint64_t)j.src_pitch};
    const cuuint32_t box[2] = {(cuuint32_t)(h.tile_rs / 8), (cuuint32_t)h.tile_rows};
    const cuuint32_t estr[2] = {1, 1};
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    CUtensorMap tm;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void*)a0, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(t_err, sizeof t_err, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return IMP_ERROR_GPU; }
    memcpy(j.tmap, &tm, 128);
    return IMP_OK;
}

int pick_variant(const ImpPass& h, const ImpJob& j) {
    // IMP_GPU_FORCE_DIRECT=1 routes everything through the general direct-from-global kernels (used by the tests to keep
    // the fallback paths — unaligned pitches, oversized footprints, large sigma — covered on the GPU).
    static const bool force_direct = [] { const char* e = getenv("IMP_GPU_FORCE_DIRECT"); return e && *e == '1'; }();
    if (force_direct) return 0;
    const int fallback = h.kind == IMP_G_CUBIC ? 3 : 0;              // cubic: the column-run kernel has no alignment requirements
    if (h.tile_smem <= 0 || !encode_tiled()) return fallback;
    if (j.src_pitch % 16) return fallback;
    const uintptr_t img = (uintptr_t)j.src;
    const uintptr_t win = img + (size_t)h.sy0 * j.src_pitch + (size_t)h.sx0 * h.sc;
    if (img % 16 && win % 16) return fallback;
    int variant = 1;                                                 // strip kernels: INTER_AREA, INTER_NN, INTER_LINEAR, index map
    if (h.kind == IMP_G_BLUR) variant = h.blur_r > 0 ? 2 : 0;        // fused blur tile kernel
    if (h.kind == IMP_G_CUBIC) variant = 4;                          // cubic tile kernel
    // the launch must fit the opt-in shared-memory limit the kernels are configured with (many LUT filters can push a
    // pass over it: ADVICE r1); the direct kernels stage only the ops
    if (variant && variant_smem(h, variant, variant_param(h, variant)) > kTileSmemLimit) return fallback;
    return variant;
}
int blur_smem_bytes(const ImpPass& h) {
    const int ops = (h.nops * (int)sizeof(ImpOp) + h.lut_bytes + 15) & ~15;
    return imp_blur_dyn_smem(h.sc, h.blur_r, ops, h.tile_rs);
}
int variant_tiles(const ImpPass& h, int variant);
int tile_stage_bytes(const ImpPass& h) {
    const int extra = h.kind == IMP_G_AREA_FRAC ? (((h.tile_ytaps * 8 + 15) & ~15) + 128) : 0;    // staged y taps + the 8 rows' int4 descriptors
    return (h.tile_smem + extra + 127) & ~127;
}
// Ring depth: as many stages as fit 72 KB (three CTAs of 72 KB + ops still share an SM's 227 KB), between 2 and 8.
// Small tiles (cfg1: 7 KB) need the depth to keep enough bytes in flight; cfg2's 23.7 KB stage gets 3 (1.5 % over 2).
int tile_stages(const ImpPass& h) {
    static const int forced = [] { const char* e = getenv("IMP_GPU_STAGES"); return e ? atoi(e) : 0; }();     // tuning knob
    if (forced >= 2 && forced <= 8) return forced;
    return std::max(2, std::min(8, (72 * 1024) / tile_stage_bytes(h)));
}
int tile_smem_bytes(const ImpPass& h, int stages) {
    const int ops = (h.nops * (int)sizeof(ImpOp) + h.lut_bytes + 15) & ~15;
    return 128 + ((ops + 127) & ~127) + stages * tile_stage_bytes(h) + 64;      // +64: padded taps past the last row
}

int variant_param(const ImpPass& h, int variant) { return variant == 1 ? tile_stages(h) : variant == 2 ? h.blur_r : 0; }
int variant_smem(const ImpPass& h, int variant, int param) {
    if (variant == 4) return imp_cubic_dyn_smem(h.sc, (h.nops * (int)sizeof(ImpOp) + h.lut_bytes + 15) & ~15, h.tile_rs, h.tile_rows);
    return variant == 1 ? tile_smem_bytes(h, param) : variant == 2 ? blur_smem_bytes(h) : ops_smem(h);
}
int variant_tiles(const ImpPass& h, int variant) {
    if (variant == 1) return (h.bw + 31) / 32;
    if (variant == 2) return ((h.bw + IMP_BLUR_TW - 1) / IMP_BLUR_TW) * ((h.bh + IMP_BLUR_TH - 1) / IMP_BLUR_TH);   // same count in destination space
    if (variant == 3) return ((h.bw + 31) / 32) * ((h.bh + 8 * IMP_CUBIC_RUN - 1) / (8 * IMP_CUBIC_RUN));
    if (variant == 4) return ((h.bw + IMP_CUBIC_T - 1) / IMP_CUBIC_T) * ((h.bh + IMP_CUBIC_T - 1) / IMP_CUBIC_T);
    return pass_tiles(h);
}

int batch_compile(imp_gpu_batch* b) {
    const int d = t_dev;
    b->dev = d;
    b->steps.clear(); b->h_jobs.clear(); b->algo_bytes = 0; b->launches = 0;
    size_t scratch = 0;
    int max_passes = 0;
    std::vector<std::vector<size_t>> off(b->items.size()), boff(b->items.size());
    for (size_t i = 0; i < b->items.size(); i++) {
        imp_gpu_plan* p = b->items[i].plan;
        int rc = plan_to_device(p);
        if (rc) return rc;
        max_passes = std::max(max_passes, (int)p->passes.size());
        b->algo_bytes += p->algo_bytes;
        scratch = plan_scratch_layout(p, scratch, off[i], boff[i], b->items[i].src, b->items[i].sp);
    }
    if (scratch > b->scratch_cap) {
        if (b->d_scratch) CK(cudaFree(b->d_scratch));
        b->d_scratch = nullptr; b->scratch_cap = 0;
        CK(cudaMalloc((void**)&b->d_scratch, scratch));
        b->scratch_cap = scratch;
    }
    // `occ`: how many CTAs of the job's shared-memory footprint fit an SM (capped at the 3 the register budget allows).
    // A launch takes the largest footprint of its group, so strip jobs are grouped by this class as well: a few big
    // tiles must not drag thousands of small ones down to two CTAs per SM.
    struct Pending { int kind, sc, variant, tmax, occ; ImpJob job; ImpPass hdr; size_t boff; };
    auto occ_class = [](const ImpPass& h, int variant, int param) {
        if (variant != 1) return 0;
        return std::max(1, std::min(3, (227 * 1024) / (variant_smem(h, variant, param) + 1024)));
    };
    for (int k = 0; k < max_passes; k++) {
        std::vector<Pending> pend;
        for (size_t i = 0; i < b->items.size(); i++) {
            const auto& it = b->items[i];
            if ((int)it.plan->passes.size() <= k) continue;
            const ImpHostPass& hp = it.plan->passes[k];
            ImpJob jb = make_job(it.plan, d, k, it.src, it.sp, it.dst, it.dp, b->d_scratch, off[i]);
            const int variant = pick_variant(hp.hdr, jb);
            if (variant == 1 || variant == 2 || variant == 4) { int rc = encode_job_tmap(hp.hdr, jb); if (rc) return rc; }
            const int param = variant_param(hp.hdr, variant);
            pend.push_back(Pending{hp.hdr.kind, hp.hdr.sc, variant, param, occ_class(hp.hdr, variant, param), jb, hp.hdr, boff[i][k]});
        }
        std::stable_sort(pend.begin(), pend.end(), [](const Pending& a, const Pending& c) {
            if (a.kind != c.kind) return a.kind < c.kind;
            if (a.sc != c.sc) return a.sc < c.sc;
            if (a.variant != c.variant) return a.variant < c.variant;
            if (a.tmax != c.tmax) return a.tmax < c.tmax;
            return a.occ < c.occ; });
        size_t s = 0;
        while (s < pend.size()) {
            size_t e = s;
            while (e < pend.size() && pend[e].kind == pend[s].kind && pend[e].sc == pend[s].sc && pend[e].variant == pend[s].variant && pend[e].tmax == pend[s].tmax && pend[e].occ == pend[s].occ) e++;
            if (pend[s].kind == IMP_G_BLUR && pend[s].variant == 0) {
                for (size_t j = s; j < e; j++) {
                    imp_gpu_batch::Step st{};
                    st.generic_blur = true; st.job = (int)b->h_jobs.size(); st.hdr = pend[j].hdr; st.smem = ops_smem(pend[j].hdr);
                    st.scratch_off = pend[j].boff;
                    b->h_jobs.push_back(pend[j].job);
                    b->steps.push_back(st); b->launches += 2;
                }
            } else {
                imp_gpu_batch::Step st{};
                st.generic_blur = false;
                st.g.kind = pend[s].kind; st.g.sc = pend[s].sc; st.g.first = (int)b->h_jobs.size(); st.g.count = (int)(e - s);
                st.g.max_tiles = 0; st.g.smem_bytes = 16; st.g.variant = pend[s].variant; st.g.tmax = pend[s].tmax;
                for (size_t j = s; j < e; j++) {
                    st.g.max_tiles = std::max(st.g.max_tiles, variant_tiles(pend[j].hdr, pend[s].variant));
                    st.g.smem_bytes = std::max(st.g.smem_bytes, variant_smem(pend[j].hdr, pend[s].variant, pend[s].tmax));
                    b->h_jobs.push_back(pend[j].job);
                }
                b->steps.push_back(st); b->launches += 1;
            }
            s = e;
        }
    }
    if (b->h_jobs.size() > b->jobs_cap) {
        if (b->d_jobs) CK(cudaFree(b->d_jobs));
        b->d_jobs = nullptr; b->jobs_cap = 0;
        CK(cudaMalloc((void**)&b->d_jobs, b->h_jobs.size() * sizeof(ImpJob)));
        b->jobs_cap = b->h_jobs.size();
    }
    if (!b->h_jobs.empty()) CK(cudaMemcpy(b->d_jobs, b->h_jobs.data(), b->h_jobs.size() * sizeof(ImpJob), cudaMemcpyHostToDevice));
    b->dirty = false;
    return IMP_OK;
}

// One frame with by-value job descriptors: no device job table, nothing to free afterwards.
// `scratch` must hold plan_scratch_layout(plan) bytes (or be null when the plan needs none).
int launch_single(imp_gpu_plan* p, const uint8_t* src, int sp, uint8_t* dst, int dp, uint8_t* scratch, cudaStream_t st) {
    std::vector<size_t> off, boff;
    plan_scratch_layout(p, 0, off, boff, src, sp);
    for (size_t k = 0; k < p->passes.size(); k++) {
        const ImpHostPass& hp = p->passes[k];
        ImpJob j = make_job(p, t_dev, k, src, sp, dst, dp, scratch, off);
        const int variant = pick_variant(hp.hdr, j);
        if (variant == 1 || variant == 2 || variant == 4) { int rc = encode_job_tmap(hp.hdr, j); if (rc) return rc; }
        if (hp.hdr.kind == IMP_G_BLUR && variant == 0) {
            CK(imp_launch_blur_generic(j, hp.hdr, (uint16_t*)(scratch + boff[k]), ops_smem(hp.hdr), st));
        } else {
            const int param = variant_param(hp.hdr, variant);
            ImpLaunchGroup g{hp.hdr.kind, hp.hdr.sc, 0, 1, variant_tiles(hp.hdr, variant), variant_smem(hp.hdr, variant, param), variant, param};
            CK(imp_launch_group(g, nullptr, &j, st));
        }
    }
    return IMP_OK;
}

size_t plan_scratch_bytes(const imp_gpu_plan* p, const uint8_t* src = nullptr, int sp = 0) { std::vector<size_t> a, b; return plan_scratch_layout(p, 0, a, b, src, sp); }

}  // namespace

extern "C" {

int imp_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int imp_gpu_init(int device) {
    if (device < 0 || device >= MAX_DEV) return fail_msg("imp_gpu_init: device index out of range");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        snprintf(t_err, sizeof t_err, "no CUDA device available (%s); libimp_gpu has no CPU fallback", e == cudaSuccess ? "count == 0" : cudaGetErrorString(e));
        return IMP_ERROR_GPU;
    }
    if (device >= n) return fail_msg("imp_gpu_init: no such device");
    CK(cudaSetDevice(device));
    {
        std::lock_guard<std::mutex> lk(g_mu);
        DevCtx& c = g_dev[device];
        if (!c.ready) {
            cudaDeviceProp prop;
            CK(cudaGetDeviceProperties(&prop, device));
            if (prop.major != 10) {
                snprintf(t_err, sizeof t_err, "device %d is sm_%d%d; libimp_gpu is built for sm_100a only", device, prop.major, prop.minor);
                return IMP_ERROR_GPU;
            }
            CK(cudaFree(0));
            CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
            CK(imp_upload_tables());
            c.ready = true;
        }
    }
    t_dev = device;
    return IMP_OK;
}

int imp_gpu_set_device(int device) {
    if (device < 0 || device >= MAX_DEV) return fail_msg("imp_gpu_set_device: device index out of range");
    if (!g_dev[device].ready) { int rc = imp_gpu_init(device); if (rc) return rc; }
    t_dev = device;
    CK(cudaSetDevice(device));
    return IMP_OK;
}

void imp_gpu_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (int d = 0; d < MAX_DEV; d++) {
        DevCtx& c = g_dev[d];
        if (!c.ready) continue;
        if (cudaSetDevice(d) == cudaSuccess) {
            cudaDeviceSynchronize();
            for (Lane& L : c.lanes) {
                if (L.d_in) cudaFree(L.d_in);
                if (L.d_out) cudaFree(L.d_out);
                if (L.d_scratch) cudaFree(L.d_scratch);
                if (L.d_stage) cudaFree(L.d_stage);
                if (L.h_in) cudaFreeHost(L.h_in);
                if (L.h_out) cudaFreeHost(L.h_out);
                if (L.st) cudaStreamDestroy(L.st);
            }
            c.lanes.clear();
            if (c.stream) cudaStreamDestroy(c.stream);
        }
        c.stream = nullptr; c.ready = false;
    }
    t_dev = -1;
}

const char* imp_gpu_last_error(void) { return t_err; }
unsigned long long imp_gpu_launch_count(void) { return imp_launches(); }

// ---- plans ---------------------------------------------------------------------------------------------
int imp_gpu_plan_create(const imp_gpu_request* req, const imp_gpu_config* cfg, int w, int h, int c, imp_gpu_plan** out, int* step) {
    if (out) *out = nullptr;
    if (!out) return IMP_ERROR_INVALID_ARGS;
    imp_gpu_plan* p = nullptr;
    try {
        p = new imp_gpu_plan();
        int rc = imp_build_plan(req, cfg, w, h, c, p, step);
        if (rc) { delete p; return rc; }
    } catch (const std::bad_alloc&) {
        delete p; return IMP_ERROR_MALLOC_FAILED;
    } catch (...) {
        delete p; return IMP_ERROR_INVALID_ARGS;
    }
    *out = p;
    return IMP_OK;
}

void imp_gpu_plan_destroy(imp_gpu_plan* plan) {
    if (!plan) return;
    { std::lock_guard<std::mutex> lk(g_mu); plan_free_device(plan); }
    if (t_dev >= 0) cudaSetDevice(t_dev);
    delete plan;
}

void imp_gpu_plan_output(const imp_gpu_plan* p, int* w, int* h, int* c) {
    if (w) *w = p->out_w; if (h) *h = p->out_h; if (c) *c = p->out_c;
}
void imp_gpu_plan_source_window(const imp_gpu_plan* p, int* x, int* y, int* w, int* h) {
    if (x) *x = p->win_x; if (y) *y = p->win_y; if (w) *w = p->win_w; if (h) *h = p->win_h;
}
int imp_gpu_plan_passes(const imp_gpu_plan* p) { return (int)p->passes.size(); }
unsigned long long imp_gpu_plan_algorithmic_bytes(const imp_gpu_plan* p) { return p->algo_bytes; }

// ---- batches -------------------------------------------------------------------------------------------
int imp_gpu_batch_create(imp_gpu_batch** out) {
    if (!out) return IMP_ERROR_INVALID_ARGS;
    *out = new (std::nothrow) imp_gpu_batch();
    return *out ? IMP_OK : IMP_ERROR_MALLOC_FAILED;
}
void imp_gpu_batch_destroy(imp_gpu_batch* b) {
    if (!b) return;
    if (b->dev >= 0 && cudaSetDevice(b->dev) == cudaSuccess) {
        if (b->d_jobs) cudaFree(b->d_jobs);
        if (b->d_scratch) cudaFree(b->d_scratch);
        if (t_dev >= 0) cudaSetDevice(t_dev);
    }
    delete b;
}
int imp_gpu_batch_clear(imp_gpu_batch* b) { b->items.clear(); b->dirty = true; return IMP_OK; }
int imp_gpu_batch_size(const imp_gpu_batch* b) { return (int)b->items.size(); }
unsigned long long imp_gpu_batch_algorithmic_bytes(const imp_gpu_batch* b) { return b->algo_bytes; }
int imp_gpu_batch_launches_per_run(const imp_gpu_batch* b) { return b->launches; }

int imp_gpu_batch_add(imp_gpu_batch* b, imp_gpu_plan* plan, const void* d_src, int sp, void* d_dst, int dp) {
    if (!b || !plan || !d_src || !d_dst) return IMP_ERROR_INVALID_ARGS;
    if (sp < plan->src_w * plan->src_c || dp < plan->out_w * plan->out_c) return IMP_ERROR_INVALID_ARGS;
    if (plan->out_c == 4 && (dp % 4 || ((uintptr_t)d_dst) % 4)) return IMP_ERROR_INVALID_ARGS;
    if (plan->src_c == 4 && (sp % 4 || ((uintptr_t)d_src) % 4)) return IMP_ERROR_INVALID_ARGS;
    try { b->items.push_back(imp_gpu_batch::Item{plan, (const uint8_t*)d_src, sp, (uint8_t*)d_dst, dp}); }
    catch (...) { return IMP_ERROR_MALLOC_FAILED; }
    b->dirty = true;
    return IMP_OK;
}

int imp_gpu_batch_launch(imp_gpu_batch* b, void* stream) {
    int rc = bind(); if (rc) return rc;
    if (!b) return IMP_ERROR_INVALID_ARGS;
    if (b->items.empty()) return IMP_OK;
    if (b->dirty || b->dev != t_dev) {
        try { rc = batch_compile(b); } catch (const std::bad_alloc&) { return IMP_ERROR_MALLOC_FAILED; }
        if (rc) return rc;
    }
    cudaStream_t st = pick_stream(stream);
    for (const auto& s : b->steps) {
        if (s.generic_blur) CK(imp_launch_blur_generic(b->h_jobs[s.job], s.hdr, (uint16_t*)(b->d_scratch + s.scratch_off), s.smem, st));
        else CK(imp_launch_group(s.g, b->d_jobs, nullptr, st));
    }
    return IMP_OK;
}

// ---- one frame -----------------------------------------------------------------------------------------
int imp_gpu_run_device(imp_gpu_plan* plan, const void* d_src, int sp, void* d_dst, int dp, void* stream) {
    int rc = bind(); if (rc) return rc;
    if (!plan || !d_src || !d_dst) return IMP_ERROR_INVALID_ARGS;
    if (sp < plan->src_w * plan->src_c || dp < plan->out_w * plan->out_c) return IMP_ERROR_INVALID_ARGS;
    if (plan->out_c == 4 && (dp % 4 || ((uintptr_t)d_dst) % 4)) return IMP_ERROR_INVALID_ARGS;
    if (plan->src_c == 4 && (sp % 4 || ((uintptr_t)d_src) % 4)) return IMP_ERROR_INVALID_ARGS;
    if ((rc = plan_to_device(plan))) return rc;
    cudaStream_t st = pick_stream(stream);
    const size_t need = plan_scratch_bytes(plan, (const uint8_t*)d_src, sp);
    uint8_t* scratch = nullptr;
    if (need) CK(cudaMallocAsync((void**)&scratch, need, st));        // stream-ordered: stays asynchronous
    rc = launch_single(plan, (const uint8_t*)d_src, sp, (uint8_t*)d_dst, dp, scratch, st);
    if (scratch) CK(cudaFreeAsync(scratch, st));
    return rc;
}

int imp_gpu_run_host(imp_gpu_plan* plan, const unsigned char* src, int src_step, unsigned char* dst, int dst_step) {
    if (!plan || !src || !dst) return IMP_ERROR_INVALID_ARGS;
    imp_gpu_plan* plans[1] = {plan};
    const unsigned char* srcs[1] = {src}; unsigned char* dsts[1] = {dst};
    int ss[1] = {src_step}, ds[1] = {dst_step};
    return imp_gpu_batch_run_host(1, plans, srcs, ss, dsts, ds, 1);
}

// End-to-end: per lane (stream) a device input/output buffer; jobs are issued round-robin so that the
// H2D of job i+1 overlaps the kernels of job i and the D2H of job i-1. Only the crop window travels.
// A host pointer that is not page-locked is first copied into a pinned staging buffer.
int imp_gpu_batch_run_host(int n, imp_gpu_plan* const* plans, const unsigned char* const* srcs, const int* src_steps,
                           unsigned char* const* dsts, const int* dst_steps, int n_streams) {
    int rc = bind(); if (rc) return rc;
    if (n <= 0) return IMP_OK;
    if (!plans || !srcs || !src_steps || !dsts || !dst_steps) return IMP_ERROR_INVALID_ARGS;
    DevCtx& ctx = g_dev[t_dev];
    std::lock_guard<std::mutex> run_lk(ctx.run_mu);
    n_streams = std::max(1, std::min(n_streams, std::min(n, 8)));
    while ((int)ctx.lanes.size() < n_streams) {
        Lane L;
        CK(cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking));
        ctx.lanes.push_back(L);
    }
    auto is_pinned = [](const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    auto finish = [&](Lane& L) -> int {
        if (L.pending < 0) return IMP_OK;
        const int i = L.pending; L.pending = -1;
        CK(cudaStreamSynchronize(L.st));
        if (L.out_staged) {
            const imp_gpu_plan* p = plans[i];
            const size_t row = (size_t)p->out_w * p->out_c;
            for (int y = 0; y < p->out_h; y++) memcpy(dsts[i] + (size_t)y * dst_steps[i], L.h_out + (size_t)y * row, row);
        }
        return IMP_OK;
    };
    auto grow = [&](uint8_t*& p, size_t& cap, size_t need, bool host) -> int {
        if (need <= cap) return IMP_OK;
        if (p) { if (host) CK(cudaFreeHost(p)); else CK(cudaFree(p)); p = nullptr; cap = 0; }
        need = (need * 5 / 4 + 4095) & ~size_t(4095);
        if (host) CK(cudaHostAlloc((void**)&p, need, cudaHostAllocDefault)); else CK(cudaMalloc((void**)&p, need));
        cap = need;
        return IMP_OK;
    };
    auto issue = [&](Lane& L, int i) -> int {
        imp_gpu_plan* p = plans[i];
        if (!p || !srcs[i] || !dsts[i]) return IMP_ERROR_INVALID_ARGS;
        int r = plan_to_device(p); if (r) return r;
        const int sc = p->src_c;
        const size_t in_row = (size_t)p->win_w * sc, out_row = (size_t)p->out_w * p->out_c;
        const int in_pitch = align16((int)in_row), out_pitch = align16((int)out_row);
        if ((r = grow(L.d_in, L.in_cap, (size_t)in_pitch * p->win_h, false))) return r;
        if ((r = grow(L.d_out, L.out_cap, (size_t)out_pitch * p->out_h, false))) return r;
        if ((r = grow(L.d_scratch, L.scratch_cap, plan_scratch_bytes(p), false))) return r;
        const uint8_t* win = srcs[i] + (size_t)p->win_y * src_steps[i] + (size_t)p->win_x * sc;
        // A 2-D host-to-device copy pays ~0.15 us per row whatever its width (measured: 3.5 KB rows move at 23 GB/s,
        // 14 KB rows at the link's 54 GB/s). Short rows therefore travel as ONE linear copy of the rows the window
        // touches (full width) and a small kernel extracts / re-pitches the window on the device.
        const uint8_t* h_lin = nullptr; size_t lin_bytes = 0; int lin_step = 0; size_t lin_off = 0;
        if (is_pinned(srcs[i])) {
            if (src_steps[i] <= 8192) {
                h_lin = srcs[i] + (size_t)p->win_y * src_steps[i]; lin_step = src_steps[i]; lin_off = (size_t)p->win_x * sc;
                lin_bytes = (size_t)(p->win_h - 1) * src_steps[i] + lin_off + in_row;
            }
        } else {
            if ((r = grow(L.h_in, L.hin_cap, in_row * p->win_h, true))) return r;
            for (int y = 0; y < p->win_h; y++) memcpy(L.h_in + (size_t)y * in_row, win + (size_t)y * src_steps[i], in_row);
            h_lin = L.h_in; lin_step = (int)in_row; lin_off = 0; lin_bytes = in_row * p->win_h;
        }
        if (h_lin && lin_step == in_pitch && lin_off == 0) {
            CK(cudaMemcpyAsync(L.d_in, h_lin, lin_bytes, cudaMemcpyHostToDevice, L.st));              // already in the device layout
        } else if (h_lin) {
            if ((r = grow(L.d_stage, L.stage_cap, lin_bytes, false))) return r;
            CK(cudaMemcpyAsync(L.d_stage, h_lin, lin_bytes, cudaMemcpyHostToDevice, L.st));
            CK(imp_launch_repitch(L.d_stage + lin_off, lin_step, L.d_in, in_pitch, (int)in_row, p->win_h, L.st));
        } else {
            CK(cudaMemcpy2DAsync(L.d_in, in_pitch, win, src_steps[i], in_row, p->win_h, cudaMemcpyHostToDevice, L.st));
        }
        // The device copy holds only the crop window; bias the base pointer so the pass's (sx0,sy0) lands on it.
        const uint8_t* biased = L.d_in - ((size_t)p->win_y * in_pitch + (size_t)p->win_x * sc);
        if ((r = launch_single(p, biased, in_pitch, L.d_out, out_pitch, L.d_scratch, L.st))) return r;
        L.out_staged = !is_pinned(dsts[i]);
        if (!L.out_staged) {
            CK(cudaMemcpy2DAsync(dsts[i], dst_steps[i], L.d_out, out_pitch, out_row, p->out_h, cudaMemcpyDeviceToHost, L.st));
        } else {
            if ((r = grow(L.h_out, L.hout_cap, out_row * p->out_h, true))) return r;
            CK(cudaMemcpy2DAsync(L.h_out, out_row, L.d_out, out_pitch, out_row, p->out_h, cudaMemcpyDeviceToHost, L.st));
        }
        L.pending = i;
        return IMP_OK;
    };
    int result = IMP_OK;
    for (int i = 0; i < n && result == IMP_OK; i++) {
        Lane& L = ctx.lanes[i % n_streams];
        if ((result = finish(L))) break;
        result = issue(L, i);
    }
    for (int s = 0; s < n_streams; s++) { int r = finish(ctx.lanes[s]); if (result == IMP_OK) result = r; }
    return result;
}

int imp_gpu_farm_run_host(int n, imp_gpu_plan* const* plans, const unsigned char* const* srcs, const int* src_steps,
                          unsigned char* const* dsts, const int* dst_steps, int n_gpus, int n_streams) {
    if (n_gpus <= 0) return IMP_ERROR_INVALID_ARGS;
    const int avail = imp_gpu_device_count();
    if (avail <= 0) return fail_msg("no CUDA device available; libimp_gpu has no CPU fallback");
    if (n_gpus > avail || n_gpus > MAX_DEV) return fail_msg("imp_gpu_farm_run_host: more GPUs requested than present");
    std::vector<int> rcs(n_gpus, IMP_OK);
    std::vector<std::string> errs(n_gpus);
    std::vector<std::thread> th;
    const int caller_dev = t_dev;
    for (int g = 0; g < n_gpus; g++) {
        th.emplace_back([&, g]() {
            int rc = imp_gpu_set_device(g);
            if (rc == IMP_OK) {
                std::vector<imp_gpu_plan*> pl; std::vector<const unsigned char*> sr; std::vector<unsigned char*> ds; std::vector<int> ss, dd;
                for (int i = g; i < n; i += n_gpus) { pl.push_back(plans[i]); sr.push_back(srcs[i]); ds.push_back(dsts[i]); ss.push_back(src_steps[i]); dd.push_back(dst_steps[i]); }
                if (!pl.empty()) rc = imp_gpu_batch_run_host((int)pl.size(), pl.data(), sr.data(), ss.data(), ds.data(), dd.data(), n_streams);
            }
            rcs[g] = rc;
            if (rc) errs[g] = t_err;
        });
    }
    for (auto& t : th) t.join();
    if (caller_dev >= 0) { t_dev = caller_dev; cudaSetDevice(caller_dev); }
    for (int g = 0; g < n_gpus; g++) if (rcs[g]) { snprintf(t_err, sizeof t_err, "gpu %d: %s", g, errs[g].c_str()); return rcs[g]; }
    return IMP_OK;
}

// ---- "next" row §8f-1: perceived brightness as a device reduction ------------------------------------------
int imp_gpu_brightness_device(const void* d_img, int pitch, int w, int h, int c, float* brightness, void* stream) {
    int rc = bind(); if (rc) return rc;
    if (!d_img || !brightness || w <= 0 || h <= 0 || (c != 1 && c != 3 && c != 4)) return IMP_ERROR_INVALID_ARGS;
    cudaStream_t st = pick_stream(stream);
    double* d_acc = nullptr;
    CK(cudaMallocAsync((void**)&d_acc, sizeof(double), st));
    CK(imp_launch_brightness((const uint8_t*)d_img, pitch, w, h, c, d_acc, st));
    double sum = 0;
    CK(cudaMemcpyAsync(&sum, d_acc, sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaFreeAsync(d_acc, st));
    CK(cudaStreamSynchronize(st));
    *brightness = (float)(sum / ((double)w * h) / 255.0);          // filters.c:728
    return IMP_OK;
}

int imp_gpu_brightness_host(const unsigned char* img, int step, int w, int h, int c, float* brightness) {
    int rc = bind(); if (rc) return rc;
    if (!img || !brightness || w <= 0 || h <= 0 || (c != 1 && c != 3 && c != 4)) return IMP_ERROR_INVALID_ARGS;
    cudaStream_t st = g_dev[t_dev].stream;
    const int pitch = align16(w * c);
    uint8_t* d = nullptr;
    CK(cudaMallocAsync((void**)&d, (size_t)pitch * h, st));
    CK(cudaMemcpy2DAsync(d, pitch, img, step, (size_t)w * c, h, cudaMemcpyHostToDevice, st));
    rc = imp_gpu_brightness_device(d, pitch, w, h, c, brightness, st);
    cudaFreeAsync(d, st);
    return rc;
}

// ---- "next" row §8f-4: ASCII (format=text) ---------------------------------------------------------------------
// Density ramps of filters.c:486-487 (the reference's data; the classic 70- and 10-level ASCII-art ramps).
static const char kAsciiWide[] = "$@B%8&WM#*oahkbdpqwmZO0QLCJUYXzcvunxrjft/\\|()1{}[]?-_+~<>i!lI;:,\"^`'. ";
static const char kAsciiNarrow[] = "@%8#*+=-:. ";

long imp_gpu_ascii_length(int width, int height) { return (long)(width + 1) * height - 1; }

int imp_gpu_ascii_host(const unsigned char* img, int step, int w, int h, int c, const char* args, unsigned char* out, long out_cap) {
    int rc = bind(); if (rc) return rc;
    if (!img || !out || w <= 0 || h <= 0 || (c != 1 && c != 3 && c != 4)) return IMP_ERROR_INVALID_ARGS;
    const long len = imp_gpu_ascii_length(w, h);
    if (out_cap < len) return IMP_ERROR_INVALID_ARGS;
    const char* table = (args && strcmp(args, "wide") == 0) ? kAsciiWide : kAsciiNarrow;     // filters.c:490-494
    const int tablelen = (int)strlen(table);
    const float factor = 256.0 / tablelen;                                                   // filters.c:496
    unsigned char lut[256];
    for (int v = 0; v < 256; v++) lut[v] = (unsigned char)table[(int)floor((double)((float)v / factor))];   // filters.c:509
    cudaStream_t st = g_dev[t_dev].stream;
    const int pitch = align16(w * c);
    uint8_t *d_img = nullptr, *d_lut = nullptr, *d_out = nullptr;
    CK(cudaMallocAsync((void**)&d_img, (size_t)pitch * h, st));
    CK(cudaMallocAsync((void**)&d_lut, 256, st));
    CK(cudaMallocAsync((void**)&d_out, (size_t)len + 1, st));
    CK(cudaMemcpy2DAsync(d_img, pitch, img, step, (size_t)w * c, h, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_lut, lut, 256, cudaMemcpyHostToDevice, st));
    CK(imp_launch_ascii(d_img, pitch, w, h, c, d_lut, d_out, st));
    CK(cudaMemcpyAsync(out, d_out, (size_t)len, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaFreeAsync(d_img, st); cudaFreeAsync(d_lut, st); cudaFreeAsync(d_out, st);
    return IMP_OK;
}

// ---- "next" row §8f-2: GIF canvas expansion ------------------------------------------------------------------
int imp_gpu_gif_expand_device(const imp_gpu_gif_frame* frames, int n, int canvas_w, int canvas_h, int destructive,
                              void* d_canvases, int canvas_pitch, void* stream) {
    int rc = bind(); if (rc) return rc;
    if (!frames || n <= 0 || canvas_w <= 0 || canvas_h <= 0 || !d_canvases || canvas_pitch < canvas_w * 4 || canvas_pitch % 4) return IMP_ERROR_INVALID_ARGS;
    cudaStream_t st = pick_stream(stream);
    // one staging buffer: [ImpGifFrame n][palettes n*1024][index planes]
    size_t idx_bytes = 0;
    for (int f = 0; f < n; f++) {
        if (!frames[f].indices || !frames[f].palette || frames[f].width <= 0 || frames[f].height <= 0 || frames[f].pitch < frames[f].width) return IMP_ERROR_INVALID_ARGS;
        idx_bytes += ((size_t)frames[f].pitch * frames[f].height + 15) & ~size_t(15);
    }
    const size_t meta_bytes = (((size_t)n * sizeof(ImpGifFrame)) + 15) & ~size_t(15), pal_bytes = (size_t)n * 1024;
    const size_t total = meta_bytes + pal_bytes + idx_bytes;
    uint8_t* h_buf = nullptr; uint8_t* d_buf = nullptr;
    CK(cudaHostAlloc((void**)&h_buf, total, cudaHostAllocDefault));
    cudaError_t e = cudaMallocAsync((void**)&d_buf, total, st);
    if (e != cudaSuccess) { cudaFreeHost(h_buf); return fail(e, "cudaMallocAsync", __LINE__); }
    ImpGifFrame* meta = reinterpret_cast<ImpGifFrame*>(h_buf);
    size_t off = meta_bytes + pal_bytes;
    for (int f = 0; f < n; f++) {
        const imp_gpu_gif_frame& g = frames[f];
        memcpy(h_buf + meta_bytes + (size_t)f * 1024, g.palette, 1024);
        memcpy(h_buf + off, g.indices, (size_t)g.pitch * g.height);
        meta[f].indices = d_buf + off; meta[f].palette = d_buf + meta_bytes + (size_t)f * 1024;
        meta[f].pitch = g.pitch; meta[f].w = g.width; meta[f].h = g.height; meta[f].left = g.left; meta[f].top = g.top;
        meta[f].dispose = g.dispose; meta[f].key = g.transparency_key; meta[f].pad_ = 0;
        off += ((size_t)g.pitch * g.height + 15) & ~size_t(15);
    }
    e = cudaMemcpyAsync(d_buf, h_buf, total, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = imp_launch_gif_expand(reinterpret_cast<const ImpGifFrame*>(d_buf), n, canvas_w, canvas_h, destructive ? 1 : 0, (uint8_t*)d_canvases, canvas_pitch, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);          // h_buf / d_buf are released below
    cudaFreeAsync(d_buf, st);
    cudaFreeHost(h_buf);
    if (e != cudaSuccess) return fail(e, "gif expand", __LINE__);
    return IMP_OK;
}

int imp_gpu_gif_expand_host(const imp_gpu_gif_frame* frames, int n, int canvas_w, int canvas_h, int destructive,
                            unsigned char* const* canvases, int canvas_step) {
    int rc = bind(); if (rc) return rc;
    if (!canvases || canvas_step < canvas_w * 4) return IMP_ERROR_INVALID_ARGS;
    cudaStream_t st = g_dev[t_dev].stream;
    const int pitch = align16(canvas_w * 4);
    uint8_t* d = nullptr;
    CK(cudaMalloc((void**)&d, (size_t)pitch * canvas_h * n));
    rc = imp_gpu_gif_expand_device(frames, n, canvas_w, canvas_h, destructive, d, pitch, st);
    for (int f = 0; f < n && rc == IMP_OK; f++) {
        cudaError_t e = cudaMemcpy2D(canvases[f], canvas_step, d + (size_t)f * pitch * canvas_h, pitch, (size_t)canvas_w * 4, canvas_h, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(e, "cudaMemcpy2D", __LINE__);
    }
    cudaFree(d);
    return rc;
}

// ---- memory helpers ------------------------------------------------------------------------------------
int imp_gpu_malloc(void** p, size_t bytes) { int rc = bind(); if (rc) return rc; CK(cudaMalloc(p, bytes)); return IMP_OK; }
int imp_gpu_free(void* p) { int rc = bind(); if (rc) return rc; CK(cudaFree(p)); return IMP_OK; }
int imp_gpu_malloc_pitch(void** p, int* pitch, int width_bytes, int height) {
    int rc = bind(); if (rc) return rc;
    const int pt = align16(width_bytes);
    CK(cudaMalloc(p, (size_t)pt * (size_t)std::max(height, 1)));
    if (pitch) *pitch = pt;
    return IMP_OK;
}
int imp_gpu_host_alloc(void** p, size_t bytes) { int rc = bind(); if (rc) return rc; CK(cudaHostAlloc(p, bytes, cudaHostAllocDefault)); return IMP_OK; }
int imp_gpu_host_free(void* p) { int rc = bind(); if (rc) return rc; CK(cudaFreeHost(p)); return IMP_OK; }
int imp_gpu_upload_2d(void* d, int dp, const void* h, int hs, int wb, int rows, void* stream) {
    int rc = bind(); if (rc) return rc;
    CK(cudaMemcpy2DAsync(d, dp, h, hs, wb, rows, cudaMemcpyHostToDevice, pick_stream(stream)));
    return IMP_OK;
}
int imp_gpu_download_2d(void* h, int hs, const void* d, int dp, int wb, int rows, void* stream) {
    int rc = bind(); if (rc) return rc;
    CK(cudaMemcpy2DAsync(h, hs, d, dp, wb, rows, cudaMemcpyDeviceToHost, pick_stream(stream)));
    return IMP_OK;
}
int imp_gpu_sync(void* stream) {
    int rc = bind(); if (rc) return rc;
    if (stream) CK(cudaStreamSynchronize((cudaStream_t)stream));
    else CK(cudaDeviceSynchronize());
    return IMP_OK;
}

}  // extern "C"
