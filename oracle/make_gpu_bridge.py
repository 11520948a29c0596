#!/usr/bin/env python
"""TEST INFRASTRUCTURE — builds oracle/_ref/libimp_ref_gpu.so: the reference with INTEGRATION.md's edits applied.

The reference's bridge.c is read where it lies (/root/reference), the handful of call-site edits that
INTEGRATION.md §2 lists are applied IN MEMORY (regex substitutions keyed on the reference's own identifiers; no
reference text is stored in this repo), the result is compiled from a temporary directory together with the
reference's untouched filters.c / helpers.c / advancedio.c and the test stubs, linked against
ngx_http_imgproc_b200/libimp_gpu.so, and the temporary source is deleted. Only the .so is kept (git-ignored,
travels to the GPU box). tests/test_gpu_parity.py then drives the reference's own RunJob through it on a B200 and
compares with the unmodified CPU build (libimp_ref.so): that is the drop-in claim, executed.

usage: python oracle/make_gpu_bridge.py            (needs /root/reference and a built libimp_gpu.so)
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("IMP_REFERENCE", "/root/reference")
PKG = os.path.join(ROOT, "ngx_http_imgproc_b200")

HELPERS = r'''
#include "imp_ops.h"
/* INTEGRATION.md §2: frames stay owned by the (stubbed) OpenCV allocator; Config -> imp_gpu_config */
static IplImage* imp_create(int w, int h, int depth, int ch) { return cvCreateImage(cvSize(w, h), depth, ch); }
static void imp_release(IplImage** im) { cvReleaseImage(im); }
static void imp_config_from(const Config* c, imp_gpu_config* g, imp_gpu_watermark* w) {
    g->max_target_w = c->MaxTargetDimensions->W;  g->max_target_h = c->MaxTargetDimensions->H;
    g->max_filters = (int)c->MaxFiltersCount;     g->allow_experiments = (int)c->AllowExperiments;
    g->watermark = NULL;
    if (c->WatermarkInfo) {
        w->pixels = c->WatermarkInfo->Pointer;    w->width = c->WatermarkInfo->Size.width;  w->height = c->WatermarkInfo->Size.height;
        w->channels = c->WatermarkInfo->Channels; w->step = c->WatermarkInfo->Step;
        w->gravity_x = c->WatermarkPosition->GravityX; w->gravity_y = c->WatermarkPosition->GravityY;
        w->offset_x = c->WatermarkPosition->OffsetX;   w->offset_y = c->WatermarkPosition->OffsetY;
        w->opacity = (int)c->WatermarkOpacity;
        g->watermark = w;
    }
}
'''

FLUSH = r'''
	{   /* INTEGRATION.md §2: one fused GPU pass per frame, before anything reads pixels */
		IplImage* imp_fr[album.Count > 0 ? album.Count : 1]; int imp_k;
		for (imp_k = 0; imp_k < album.Count; imp_k++) imp_fr[imp_k] = album.Frames[imp_k].Image;
		answer->Code = imp_FlushAll(imp_fr, album.Count);
		for (imp_k = 0; imp_k < album.Count; imp_k++) album.Frames[imp_k].Image = imp_fr[imp_k];
		if (answer->Code) { goto finalize; }
	}
'''


def sub_once(pattern, repl, text, what, flags=0):
    new, n = re.subn(pattern, repl, text, count=1, flags=flags)
    if n != 1:
        raise SystemExit(f"make_gpu_bridge: edit '{what}' did not apply (reference layout changed?)")
    return new


def patched_bridge(src: str) -> str:
    s = src
    s = sub_once(r'(#include "advancedio.h"\n)', lambda m: m.group(1) + HELPERS, s, "helpers")
    s = sub_once(r'(void OnEnvStart\(\) \{\n)[^\n]*\n', lambda m: m.group(1) + "\timp_gpu_init(0); imp_ops_set_image_allocator(imp_create, imp_release);\n", s, "OnEnvStart")
    s = sub_once(r'(\tanswer->Step = IMP_STEP_CROP;\n)', lambda m: "\timp_gpu_config gcfg; imp_gpu_watermark gwm; imp_config_from(config, &gcfg, &gwm);\n" + m.group(1), s, "config")
    s = sub_once(r'answer->Code = Crop\(&image, crop, gravity\);', "answer->Code = imp_Crop(&image, crop, gravity);", s, "Crop")
    s = sub_once(r'answer->Code = Resize\(&image, resize, config, simple\);', "answer->Code = imp_Resize(&image, resize, &gcfg, simple);", s, "Resize")
    s = sub_once(r'\t\tif \(image->nChannels == 1\) \{\n(?:[^\n]*\n){4}\t\t\}\n', "", s, "gray->BGR block")
    s = sub_once(r'answer->Code = Filter\(&image, filters\[i\], config->AllowExperiments\);', "answer->Code = imp_Filter(&image, filters[i], config->AllowExperiments);", s, "Filter")
    s = sub_once(r'answer->Code = Watermark\(image, config\);', "answer->Code = imp_Watermark(image, &gcfg);", s, "Watermark")
    s = sub_once(r'\t\t\tBlendWithPaper\(image\);', "\t\t\timp_BlendWithPaper(image);", s, "BlendWithPaper")
    s = sub_once(r'(\t// alternative exit points\n)', lambda m: FLUSH + m.group(1), s, "flush")
    s = sub_once(r'(finalize:.*?)\t\t\t\tcvReleaseImage\(&image\);', lambda m: m.group(1) + "\t\t\t\timp_Discard(image); cvReleaseImage(&image);", s, "finalize", flags=re.S)
    return s


def main():
    lib = os.path.join(PKG, "libimp_gpu.so")
    if not os.path.exists(os.path.join(REF, "bridge.c")):
        print("make_gpu_bridge:", REF, "absent, keeping prebuilt oracle/_ref/libimp_ref_gpu.so (if any)")
        return 0
    if not os.path.exists(lib):
        raise SystemExit("make_gpu_bridge: build ngx_http_imgproc_b200/libimp_gpu.so first")
    out_dir = os.path.join(HERE, "_ref")
    os.makedirs(out_dir, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="imp_gpu_bridge_")
    try:
        with open(os.path.join(REF, "bridge.c")) as f:
            patched = patched_bridge(f.read())
        gen = os.path.join(tmp, "bridge_gpu.c")
        with open(gen, "w") as f:
            f.write(patched)
        cmd = [os.environ.get("CC", "gcc"), "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-w",
               "-I", os.path.join(HERE, "shim"), "-I", REF, "-I", os.path.join(ROOT, "include"),
               "-o", os.path.join(out_dir, "libimp_ref_gpu.so"),
               os.path.join(REF, "filters.c"), os.path.join(REF, "helpers.c"), gen, os.path.join(REF, "advancedio.c"),
               os.path.join(HERE, "ref_stubs.c"), os.path.join(HERE, "fake_freeimage.c"), os.path.join(HERE, "imp_oracle.c"),
               "-L", PKG, "-limp_gpu", "-Wl,-rpath,$ORIGIN/../../ngx_http_imgproc_b200", "-lm"]
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print(os.path.join(out_dir, "libimp_ref_gpu.so"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
