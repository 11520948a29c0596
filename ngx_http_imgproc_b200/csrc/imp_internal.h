// imp_internal.h — host-side internals of libimp_gpu.so (not part of the ABI).
#pragma once
#include <stdint.h>
#include <atomic>
#include <memory>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "imp_plan.h"
#include "../../include/imp_gpu.h"

// ---- planner output (host) ------------------------------------------------------------------------
struct ImpHostPass {
    ImpPass hdr;                       // offsets filled by finalize()
    std::vector<uint8_t> blob;         // ImpPass + tables + ops + LUTs, ready to upload
    int in_w, in_h, in_c;              // input image of the pass (whole image, before the window)
    int out_w, out_h, out_c;           // stored image
    int uses_watermark;
    double sigma;                      // BLUR only
};

// A decoded overlay (what PrepareWatermark leaves in the conf pool, bridge.c:199-237). The reference decodes it ONCE at
// configuration time and every request blends the same pixels (bridge.c:239-281); here it is interned by content, shared
// by every plan that uses it and uploaded once per device (imp_gpu_upload_watermark, or lazily on first use).
struct ImpWmImage {
    std::vector<uint8_t> pixels;       // tightly packed, w*c bytes per row
    int w = 0, h = 0, c = 0;
    unsigned long long hash = 0;
    struct Dev { uint8_t* d = nullptr; int pitch = 0; };
    Dev dev[16];
    ~ImpWmImage();                     // releases the device copies through imp_wm_dev_release (set by imp_gpu.cu)
};
extern void (*imp_wm_dev_release)(ImpWmImage*);
// imp_planner.cpp: the shared image holding exactly these pixels (registered on first sight; bounded registry).
std::shared_ptr<ImpWmImage> imp_wm_intern(const imp_gpu_watermark* wm);
void imp_wm_registry_clear();

struct imp_gpu_plan {
    std::vector<ImpHostPass> passes;
    int src_w, src_h, src_c;
    int win_x, win_y, win_w, win_h;    // crop window in the source
    int out_w, out_h, out_c;
    unsigned long long algo_bytes;
    std::shared_ptr<ImpWmImage> wm;    // overlay the plan blends (shared, uploaded once per device), or null
    std::atomic<int> refs{1};          // the plan cache and every imp_gpu_plan_create caller hold one reference each
    // per-device state: ONE stream-ordered allocation holds every pass blob (+ vignette tables), filled by an
    // asynchronous copy from pinned staging; `ready_ev` orders the first kernels behind it
    struct Dev {
        std::vector<uint8_t*> pass_blobs;          // pointers into `arena`
        uint8_t* arena = nullptr;
        cudaEvent_t ready_ev = nullptr;
        std::atomic<bool> settled{false};          // the upload is known to have completed
        bool ready = false;
    };
    Dev dev[16];
};

// imp_planner.cpp. `dry`: validate only — same codes, steps and output geometry, but no tables, LUTs, blobs or
// watermark pixels are materialised (what each recorded operator of the imp_ops layer needs).
int imp_build_plan(const imp_gpu_request* req, const imp_gpu_config* cfg, int w, int h, int c,
                   imp_gpu_plan* plan, int* step, bool dry = false);

// imp_kernels.cu
#define IMP_CUBIC_RUN 8           // imp_cubic_run_kernel: output rows per thread; a CTA covers 32 x (8 * IMP_CUBIC_RUN) pixels
struct ImpLaunchGroup {
    int kind, sc;                // kernel variant
    int first, count;            // jobs [first, first+count) of the device job table
    int max_tiles;               // largest tile count of any job in the group (grid.x)
    int smem_bytes;              // ops + LUT staging (+ source tile for tile variants)
    int variant;                 // 0 = direct-from-global kernel, 1 = shared-memory tile kernel (imp_tiles.cuh)
    int tmax;                    // tile variant: number of ring stages (2..8)
    int light;                   // gather / cubic tile kernels: the table-ops-only instantiation (ImpPass::light of every job)
};
// d_jobs == nullptr: a single job passed by value (`one`), no device job table needed.
cudaError_t imp_launch_group(const ImpLaunchGroup& g, const ImpJob* d_jobs, const ImpJob* one, cudaStream_t st);
// Two-kernel Gaussian through a u16 scratch (any sigma). One job, passed by value.
cudaError_t imp_launch_blur_generic(const ImpJob& job, const ImpPass& hdr, uint16_t* d_scratch, int smem_bytes, cudaStream_t st);
unsigned long long imp_launches();
void imp_count_launches(int n);
cudaError_t imp_upload_tables();
// one translation unit per kernel family (each with its own copy of the per-byte division tables, imp_pixel.cuh)
cudaError_t imp_upload_tables_strip();
cudaError_t imp_upload_tables_blur();
cudaError_t imp_upload_tables_cubic();
cudaError_t imp_upload_tables_gather();
unsigned imp_debug_flags_strip(); unsigned imp_debug_flags_blur(); unsigned imp_debug_flags_cubic(); unsigned imp_debug_flags_gather();
cudaError_t imp_launch_strip(const ImpLaunchGroup& g, const ImpJob* d_jobs, const ImpJob* one, cudaStream_t st);        // imp_k_strip.cu
cudaError_t imp_launch_blur_tile(const ImpLaunchGroup& g, const ImpJob* d_jobs, const ImpJob* one, cudaStream_t st);    // imp_k_blur.cu
cudaError_t imp_launch_cubic_tile(const ImpLaunchGroup& g, const ImpJob* d_jobs, const ImpJob* one, cudaStream_t st);   // imp_k_cubic.cu
cudaError_t imp_launch_gather_tile(const ImpLaunchGroup& g, const ImpJob* d_jobs, const ImpJob* one, cudaStream_t st);  // imp_k_gather.cu
// rows of `row_bytes` bytes from (src, sp) to (dst, dp); dst and dp are multiples of 4 (the library's own pitched buffers)
cudaError_t imp_launch_repitch(const uint8_t* d_src, int sp, uint8_t* d_dst, int dp, int row_bytes, int rows, cudaStream_t st);
cudaError_t imp_launch_gif_expand(const ImpGifFrame* d_frames, int n, int cw, int ch, int destructive, uint8_t* d_canvases, int cpitch, cudaStream_t st);
cudaError_t imp_launch_ascii(const uint8_t* d_img, int pitch, int w, int h, int c, const uint8_t* d_lut, uint8_t* d_out, cudaStream_t st);
cudaError_t imp_launch_brightness(const uint8_t* d_img, int pitch, int w, int h, int c, double* d_acc, cudaStream_t st);
cudaError_t imp_build_vignette_table(float* d_tab, int nx, int ny, float maxr, float intensity, cudaStream_t st);          // per-device constant tables of imp_pixel.cuh
