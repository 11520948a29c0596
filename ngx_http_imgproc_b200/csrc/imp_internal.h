// imp_internal.h — host-side internals of libimp_gpu.so (not part of the ABI).
#pragma once
#include <stdint.h>
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

struct imp_gpu_plan {
    std::vector<ImpHostPass> passes;
    int src_w, src_h, src_c;
    int win_x, win_y, win_w, win_h;    // crop window in the source
    int out_w, out_h, out_c;
    unsigned long long algo_bytes;
    // watermark (host copy, tightly packed) — uploaded lazily per device
    std::vector<uint8_t> wm_pixels;
    int wm_w = 0, wm_h = 0, wm_c = 0;
    // per-device state
    struct Dev {
        std::vector<uint8_t*> pass_blobs;
        uint8_t* wm = nullptr; int wm_pitch = 0;
        std::vector<float*> vignette_tabs;         // one per vignette op that is tabulated
        bool ready = false;
    };
    Dev dev[16];
};

// imp_planner.cpp
int imp_build_plan(const imp_gpu_request* req, const imp_gpu_config* cfg, int w, int h, int c,
                   imp_gpu_plan* plan, int* step);

// imp_kernels.cu
#define IMP_CUBIC_RUN 8           // imp_cubic_run_kernel: output rows per thread; a CTA covers 32 x (8 * IMP_CUBIC_RUN) pixels
struct ImpLaunchGroup {
    int kind, sc;                // kernel variant
    int first, count;            // jobs [first, first+count) of the device job table
    int max_tiles;               // largest tile count of any job in the group (grid.x)
    int smem_bytes;              // ops + LUT staging (+ source tile for tile variants)
    int variant;                 // 0 = direct-from-global kernel, 1 = shared-memory tile kernel (imp_tiles.cuh)
    int tmax;                    // tile variant: number of ring stages (2..8)
};
// d_jobs == nullptr: a single job passed by value (`one`), no device job table needed.
cudaError_t imp_launch_group(const ImpLaunchGroup& g, const ImpJob* d_jobs, const ImpJob* one, cudaStream_t st);
// Two-kernel Gaussian through a u16 scratch (any sigma). One job, passed by value.
cudaError_t imp_launch_blur_generic(const ImpJob& job, const ImpPass& hdr, uint16_t* d_scratch, int smem_bytes, cudaStream_t st);
unsigned long long imp_launches();
cudaError_t imp_upload_tables();
// rows of `row_bytes` bytes from (src, sp) to (dst, dp); dst and dp are multiples of 4 (the library's own pitched buffers)
cudaError_t imp_launch_repitch(const uint8_t* d_src, int sp, uint8_t* d_dst, int dp, int row_bytes, int rows, cudaStream_t st);
cudaError_t imp_launch_gif_expand(const ImpGifFrame* d_frames, int n, int cw, int ch, int destructive, uint8_t* d_canvases, int cpitch, cudaStream_t st);
cudaError_t imp_launch_ascii(const uint8_t* d_img, int pitch, int w, int h, int c, const uint8_t* d_lut, uint8_t* d_out, cudaStream_t st);
cudaError_t imp_launch_brightness(const uint8_t* d_img, int pitch, int w, int h, int c, double* d_acc, cudaStream_t st);
cudaError_t imp_build_vignette_table(float* d_tab, int n, float maxr, float intensity, cudaStream_t st);          // per-device constant tables of imp_pixel.cuh
