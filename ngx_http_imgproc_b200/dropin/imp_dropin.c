/*
 * imp_dropin.c — the reference's OWN operator symbols (bridge.h:2-7, filters.h:1-16,30) implemented on libimp_gpu.so.
 *
 * This file is compiled INSIDE the nginx module, next to bridge.c (add it to NGX_ADDON_SRCS in the module's `config`), so
 * it sees the module's real required.h — Config, Album, IplImage come from there, nothing is re-declared. With it linked
 * in, the definitions it replaces are deleted from the reference and every call site stays as it is:
 *
 *   bridge.c   OnEnvStart, OnEnvDestroy (:10-16), Crop (:18-141), Resize (:143-197), Watermark (:239-281)
 *   filters.c  the 14 filter callbacks Flip .. Scanline (:72-455) and BlendWithPaper (:666-687)
 *
 * filters.c's Filter / CallbackMap / CheckDestructive (:5-70) stay untouched and now dispatch to the callbacks below;
 * ASCII, CalcPerceivedBrightness and helpers.c stay on the host. RunJob itself gains ONE statement before its
 * "alternative exit points" (bridge.c:658):   answer->Code = imp_FlushAlbum(&album); if (answer->Code) goto finalize;
 * and its gray->BGR block (bridge.c:613-618) goes away. Every function here validates its arguments
 * exactly like the one it replaces (same IMP_* code, image untouched on error), records the operation and fixes the
 * IplImage header up; imp_FlushAlbum runs what was recorded for all frames as fused GPU passes.
 */
#include "required.h"
#include "bridge.h"
#include "filters.h"
#include "imp_ops.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* frames stay owned by OpenCV's allocator: RunJob releases them with cvReleaseImage (bridge.c:714-722) */
static IplImage* imp_create(int w, int h, int depth, int ch) { return cvCreateImage(cvSize(w, h), depth, ch); }
static void      imp_release(IplImage** im)                  { cvReleaseImage(im); }

/* Config (required.h:110-120) -> the plain struct of imp_gpu.h; the watermark pixels stay in the conf pool
 * (PrepareWatermark, bridge.c:199-237) and are interned + uploaded once per worker by the library */
static void imp_config_from(const Config* c, imp_gpu_config* g, imp_gpu_watermark* w) {
    g->max_target_w = c->MaxTargetDimensions ? c->MaxTargetDimensions->W : 0;
    g->max_target_h = c->MaxTargetDimensions ? c->MaxTargetDimensions->H : 0;
    g->max_filters = (int)c->MaxFiltersCount;
    g->allow_experiments = (int)c->AllowExperiments;
    g->watermark = NULL;
    if (c->WatermarkInfo) {
        w->pixels = c->WatermarkInfo->Pointer;
        w->width = c->WatermarkInfo->Size.width;  w->height = c->WatermarkInfo->Size.height;
        w->channels = c->WatermarkInfo->Channels; w->step = c->WatermarkInfo->Step;
        w->gravity_x = c->WatermarkPosition->GravityX; w->gravity_y = c->WatermarkPosition->GravityY;
        w->offset_x = c->WatermarkPosition->OffsetX;   w->offset_y = c->WatermarkPosition->OffsetY;
        w->opacity = (int)c->WatermarkOpacity;
        g->watermark = w;
    }
}

/* module.c:100-107 calls these per worker process, i.e. after fork: the CUDA context is created here */
void OnEnvStart()   { imp_gpu_init(0); imp_ops_set_image_allocator(imp_create, imp_release); }
void OnEnvDestroy() { imp_gpu_shutdown(); }

int Crop(IplImage** pointer, char* args, char* gravity) { return imp_Crop(pointer, args, gravity); }

int Resize(IplImage** pointer, char* args, Config* config, int simple) {
    imp_gpu_config g; imp_gpu_watermark w;
    imp_config_from(config, &g, &w);
    return imp_Resize(pointer, args, &g, simple);
}

int Watermark(IplImage* image, Config* config) {
    imp_gpu_config g; imp_gpu_watermark w;
    imp_config_from(config, &g, &w);
    return imp_Watermark(image, &g);
}

void BlendWithPaper(IplImage* source) { imp_BlendWithPaper(source); }

/* The callbacks filters.c's CallbackMap points at. Filter (filters.c:43-70, unchanged) has already split "name=args",
 * matched the name and checked the experimental flag, so the operation is recorded as allowed. */
static int imp_record(IplImage** pointer, const char* name, char* args) {
    size_t n = strlen(name) + 1 + strlen(args) + 1;
    char* request = (char*)malloc(n);
    int rc;
    if (!request) return IMP_ERROR_MALLOC_FAILED;
    snprintf(request, n, "%s=%s", name, args);
    rc = imp_Filter(pointer, request, 1);
    free(request);
    return rc;
}
int Flip    (IplImage** pointer, char* args) { return imp_record(pointer, "flip", args); }
int Rotate  (IplImage** pointer, char* args) { return imp_record(pointer, "rotate", args); }
int Modulate(IplImage** pointer, char* args) { return imp_record(pointer, "modulate", args); }
int Colorize(IplImage** pointer, char* args) { return imp_record(pointer, "colorize", args); }
int Blur    (IplImage** pointer, char* args) { return imp_record(pointer, "blur", args); }
int Gamma   (IplImage** pointer, char* args) { return imp_record(pointer, "gamma", args); }
int Contrast(IplImage** pointer, char* args) { return imp_record(pointer, "contrast", args); }
int Gradmap (IplImage** pointer, char* args) { return imp_record(pointer, "gradmap", args); }
int Vignette(IplImage** pointer, char* args) { return imp_record(pointer, "vignette", args); }
int Gotham  (IplImage** pointer, char* args) { return imp_record(pointer, "gotham", args); }
int Lomo    (IplImage** pointer, char* args) { return imp_record(pointer, "lomo", args); }
int Kelvin  (IplImage** pointer, char* args) { return imp_record(pointer, "kelvin", args); }
int Rainbow (IplImage** pointer, char* args) { return imp_record(pointer, "rainbow", args); }
int Scanline(IplImage** pointer, char* args) { return imp_record(pointer, "scanline", args); }

/* Runs everything recorded for the album's frames (bridge.c:576-656 loops) as fused GPU passes: one batched call. */
int imp_FlushAlbum(Album* album) {
    IplImage* stack[64];
    IplImage** fr = stack;
    int k, rc;
    if (album->Count <= 0) return IMP_OK;
    if (album->Count > 64) { fr = (IplImage**)malloc(sizeof(IplImage*) * (size_t)album->Count); if (!fr) return IMP_ERROR_MALLOC_FAILED; }
    for (k = 0; k < album->Count; k++) fr[k] = album->Frames[k].Image;
    rc = imp_FlushAll(fr, album->Count);
    for (k = 0; k < album->Count; k++) album->Frames[k].Image = fr[k];
    if (fr != stack) free(fr);
    return rc;
}
