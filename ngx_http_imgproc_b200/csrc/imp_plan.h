// imp_plan.h — structures shared by the host planner and the device kernels.
//
// A request (crop / resize / filter list / watermark / flatten; RunJob steps 3-7, bridge.c:574-656) is
// lowered on the host into 1..n PASSES. Pass 0 gathers from the decoded frame (crop folded into the
// source addressing, resize as the gather); every Gaussian blur (filters.c:192-207) starts a new pass
// whose gather is the stencil. All other filters, the watermark and the paper-flatten are pointwise
// OPS executed in registers between a pass's gather and its store, so a blur-free request is a single
// kernel and each frame makes one HBM round trip.
//
// Geometry filters (flip, rotate; filters.c:72-133) never move pixels by themselves: all passes work
// in the orientation of the BASE frame (the frame right after crop+resize). Each coordinate-dependent
// op carries the map base(x,y) -> its own frame, and the last pass applies the accumulated map when
// storing. This is exact because the Gaussian kernel is symmetric and sigma_x == sigma_y, so blur
// commutes with the eight flips/transposes.
#pragma once
#include <stdint.h>

enum ImpGather : int {
    IMP_G_COPY = 0,       // index map only (crop / no resize)
    IMP_G_NN,             // cvResize CV_INTER_NN        
    IMP_G_AREA_INT,       // cvResize CV_INTER_AREA, integer scales (A.2)
    IMP_G_AREA_FRAC,      // cvResize CV_INTER_AREA, generic        (A.3)
    IMP_G_CUBIC,          // cvResize CV_INTER_CUBIC     (A.4)
    IMP_G_LINEAR,         // cv::INTER_LINEAR extension  (A.4)
    IMP_G_BLUR,           // cvSmooth CV_GAUSSIAN        (A.5)
    IMP_G_COUNT
};

enum ImpOpKind : int {
    IMP_OP_MODULATE = 1,  // ModulateHSV filters.c:524-547
    IMP_OP_ADDCOLOR,      // AlphaBlendAddColor filters.c:608-616
    IMP_OP_LUT_ALL,       // ApplyGamma filters.c:549-559 (every channel, alpha too)
    IMP_OP_CONTRAST,      // BrightnessContrast filters.c:595-605
    IMP_OP_GRADMAP,       // Gradmap filters.c:260-277
    IMP_OP_VIGNETTE,      // Vignette filters.c:295-323 + RadialGradient :693-703
    IMP_OP_LOMO,          // Lomo filters.c:335-346
    IMP_OP_RAINBOW,       // Rainbow filters.c:356-403
    IMP_OP_SCANLINE,      // Scanline filters.c:405-455
    IMP_OP_WATERMARK,     // Watermark bridge.c:239-281 + AlphaBlendOver filters.c:619-662
    IMP_OP_PAPER,         // BlendWithPaper filters.c:666-687
};

// base-frame (x,y) -> coordinates in another frame of size w x h:
//   (u,v) = swap ? (y,x) : (x,y);  X = flipx ? w-1-u : u;  Y = flipy ? h-1-v : v
struct ImpFrameMap {
    int swap, flipx, flipy, w, h;
};

struct ImpOp {                // 64 bytes
    int   kind;
    int   i[6];
    float f[4];
    ImpFrameMap map;          // frame the op sees (vignette, scanline, watermark)
};

#define IMP_MAX_OPS 48

// One pass, as the kernels see it. Lives at offset 0 of the pass blob in device memory; every *_off
// is a byte offset into that blob (0 = absent).
struct ImpPass {
    int kind;                 // ImpGather
    int sc;                   // channels of the pass input (1, 3 or 4)
    int oc;                   // channels the ops and the store see (3 or 4; 1 only when sc==1 and gray is kept)
    int sx0, sy0;             // origin of the source window inside the input image (crop), pixels
    int sw, sh;               // size of the source window == clamp bounds of the gather
    int bw, bh;               // base-frame size == the gather's output size
    int nx, ny;               // AREA_INT block size
    int ksize;                // CUBIC 4 / LINEAR 2 ; BLUR: tap count n
    int simd_end;             // CUBIC: bytes of a row handled by OpenCV's float path (A.4)
    float area_scale;         // AREA_INT: f32(1/(nx*ny))
    int xofs_off, xcoef_off;  // NN: xofs[bw]; CUBIC/LINEAR: xofs[bw], short xcoef[bw*ksize];
    int yofs_off, ycoef_off;  // AREA_FRAC: int2 range[b] (first tap, count) + coef = {int si; float a}[taps]
    int taps_off;             // BLUR: int taps[n]; CUBIC: float ycoef[bh*4] = short coefficient * 2^-22 (exact)
    int blur_r;               // BLUR tile kernel: padded tap radius (3, 6, 9 or 12); 0 = generic two-launch path only
    int tapsr_off;            // BLUR tile kernel: int taps[2*blur_r+1] (zero taps trimmed, then zero-padded symmetrically)
    int max_xtaps, max_ytaps; // AREA_FRAC: largest tap count per output column / row
    int tile_rs, tile_rows;   // tile kernels: shared-memory row stride (bytes) and max source rows per 32x8 tile
    int tile_smem;            // tile kernels: source-tile bytes (tile_rs*tile_rows); 0 = no tile variant for this pass
    int xtile_off, ytile_off; // tile kernels: int2 per tile column {first px, last px}; int4 per tile row {first src row, rows, first y tap, y taps}
    int tile_ytaps;           // strip kernel: y-tap entries staged per tile (max over tiles, incl. alignment slack)
    int yrow4_off;            // strip kernel: int4 per output row {byte offset of its first source row inside the tile, y taps, index of its first tap in the tile's staged taps, 0}
    int nops;
    int ops_off;              // ImpOp[nops]
    int lut_off, lut_bytes;   // LUT area (gamma 256 B each, gradmap 768 B each)
    ImpFrameMap out;          // base (x,y) -> destination pixel
    int dc;                   // destination channels (== oc)
    int blob_bytes;
};

// One job of a batch: which pass blob, where the pixels are.
struct ImpJob {
    const uint8_t* src;       // input image of this pass (top-left of the whole image, not of the window)
    uint8_t*       dst;
    const uint8_t* pass;      // device pointer to the ImpPass blob
    const uint8_t* wm;        // watermark pixels on this device (or null)
    int src_pitch, dst_pitch, wm_pitch, wm_c;
    int tm_x0;                // strip kernels: byte offset of the source window's first pixel inside the tensor map's row
    int pad_[15];
    // strip kernels: CUtensorMap (2-D, 8-byte elements) over the 16-byte aligned source window of this job, so the
    // TMA engine fetches a whole tile (box = tile_rs bytes x tile_rows rows) with ONE instruction.
    alignas(64) unsigned char tmap[128];
};

// One GIF frame as the producer of the hot path sees it (advancedio.c:126-186): palette indices in FreeImage scanline
// order (bottom-up), placement on the canvas, disposal and transparency key, palette as RGBQUAD (B,G,R,x).
struct ImpGifFrame {
    const uint8_t* indices;   // device pointer
    const uint8_t* palette;   // device pointer, 256 * 4 bytes
    int pitch, w, h, left, top, dispose, key, pad_;
};
