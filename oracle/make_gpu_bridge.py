#!/usr/bin/env python
"""TEST INFRASTRUCTURE — builds oracle/_ref/libimp_ref_gpu.so: the reference with INTEGRATION.md's edits applied.

The reference's bridge.c, filters.c and advancedio.c are read where they lie (/root/reference); the definitions that
ngx_http_imgproc_b200/dropin/imp_dropin.c replaces under their own names are deleted IN MEMORY (brace-matched on the
reference's own identifiers; no reference text is stored in this repo), RunJob gets INTEGRATION.md §2's one flush
statement, LoadGIF's per-pixel canvas loop becomes INTEGRATION.md §4's one call, and the result is compiled from a
temporary directory together with imp_dropin.c, the reference's untouched helpers.c and the test stubs, linked against ngx_http_imgproc_b200/libimp_gpu.so; the temporary sources are
deleted. RunJob's operator call sites (Crop, Resize, Filter, Watermark, BlendWithPaper) are NOT edited. Only the .so is kept (git-ignored,
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

DROPIN = os.path.join(PKG, "dropin", "imp_dropin.c")

# the definitions ngx_http_imgproc_b200/dropin/imp_dropin.c replaces (its header comment lists them with file:line)
BRIDGE_REPLACED = ["OnEnvStart", "OnEnvDestroy", "Crop", "Resize", "Watermark"]
FILTERS_REPLACED = ["Flip", "Rotate", "Modulate", "Colorize", "Blur", "Gamma", "Contrast", "Gradmap", "Vignette", "Gotham", "Lomo",
                    "Kelvin", "Rainbow", "Scanline", "BlendWithPaper"]

FLUSH = "\tanswer->Code = imp_FlushAlbum(&album); if (answer->Code) { goto finalize; }   /* INTEGRATION.md §2: the one added statement */\n"


def sub_once(pattern, repl, text, what, flags=0):
    new, n = re.subn(pattern, repl, text, count=1, flags=flags)
    if n != 1:
        raise SystemExit(f"make_gpu_bridge: edit '{what}' did not apply (reference layout changed?)")
    return new


def delete_function(text: str, name: str) -> str:
    """Removes the top-level definition `<type> name(...) { ... }` (brace-matched; the reference's sources hold no braces
    in strings or comments inside these bodies — the compile that follows would catch a wrong cut)."""
    m = re.search(r'^(?:int|void)\s+' + re.escape(name) + r'\s*\([^;{]*\)\s*\{', text, flags=re.M)
    if not m:
        raise SystemExit(f"make_gpu_bridge: definition of {name} not found (reference layout changed?)")
    depth, i = 1, m.end()
    while depth and i < len(text):
        depth += {'{': 1, '}': -1}.get(text[i], 0)
        i += 1
    if depth:
        raise SystemExit(f"make_gpu_bridge: unbalanced braces after {name}")
    return text[:m.start()] + f"/* {name}: provided by imp_dropin.c */\n" + text[i:]


def patched_bridge(src: str) -> str:
    """bridge.c with INTEGRATION.md §2 applied: five replaced definitions deleted, the gray->BGR block deleted, one flush
    statement and one imp_Discard added. Every operator CALL SITE of RunJob stays as the reference wrote it."""
    s = src
    for name in BRIDGE_REPLACED:
        s = delete_function(s, name)
    s = sub_once(r'(#include "advancedio.h"\n)', lambda m: m.group(1) + '#include "imp_ops.h"\nint imp_FlushAlbum(Album* album);\n', s, "include")
    s = sub_once(r'\t\tif \(image->nChannels == 1\) \{\n(?:[^\n]*\n){4}\t\t\}\n', "", s, "gray->BGR block")
    s = sub_once(r'(\t// alternative exit points\n)', lambda m: FLUSH + m.group(1), s, "flush")
    s = sub_once(r'(finalize:.*?)\t\t\t\tcvReleaseImage\(&image\);', lambda m: m.group(1) + "\t\t\t\timp_Discard(image); cvReleaseImage(&image);", s, "finalize", flags=re.S)
    for call in ("Crop(&image, crop, gravity)", "Resize(&image, resize, config, simple)", "Filter(&image, filters[i], config->AllowExperiments)",
                 "Watermark(image, config)", "BlendWithPaper(image)"):
        if "answer->Code = " + call not in s and "\t" + call + ";" not in s:
            raise SystemExit(f"make_gpu_bridge: call site '{call}' is no longer in RunJob")
    return s


def patched_filters(src: str) -> str:
    """filters.c without the 14 callbacks and BlendWithPaper; Filter, CallbackMap, CheckDestructive, ASCII,
    CalcPerceivedBrightness and the pixel helpers stay as they are."""
    s = src
    for name in FILTERS_REPLACED:
        s = delete_function(s, name)
    return s


GIF_PAGE = ("        result->Error = imp_AlbumGifPage(result, frameid, isdestructive, FreeImage_GetBits(frame), FreeImage_GetPitch(frame), "
            "w, h, left, top, palette);   /* INTEGRATION.md §4 */\n        if (result->Error) { return; }\n")


def patched_advancedio(src: str) -> str:
    """advancedio.c with INTEGRATION.md §4 applied: LoadGIF's per-pixel canvas loop (the `int x, y;` declaration and the
    `for (y ...)` nest that follows it) is replaced by ONE call that hands the page to the drop-in; everything else —
    page walking, tags, palette, 8-bit conversion, frame creation, unlocking, the `page` tail — stays as the reference wrote it."""
    start = src.index("static void LoadGIF(")
    m = re.compile(r'^[ \t]*int x, y;\s*\n[ \t]*for \(y = 0; y < canvasH; y\+\+\) \{', re.M).search(src, start)
    if not m:
        raise SystemExit("make_gpu_bridge: LoadGIF's canvas loop not found (reference layout changed?)")
    depth, i = 1, m.end()
    while depth and i < len(src):
        depth += {'{': 1, '}': -1}.get(src[i], 0)
        i += 1
    if depth:
        raise SystemExit("make_gpu_bridge: unbalanced braces in LoadGIF")
    s = src[:m.start()] + GIF_PAGE + src[i:]
    proto = ("int imp_AlbumGifPage(Album* album, int frameid, int isdestructive, const unsigned char* bits, int pitch, int width, "
             "int height, int left, int top, const void* palette);\n")
    return sub_once(r'(static void LoadGIF\()', lambda mm: proto + mm.group(1), s, "LoadGIF prototype")


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
        gen = {}
        for name, fn in (("bridge.c", patched_bridge), ("filters.c", patched_filters), ("advancedio.c", patched_advancedio)):
            with open(os.path.join(REF, name)) as f:
                text = fn(f.read())
            gen[name] = os.path.join(tmp, name.replace(".c", "_gpu.c"))
            with open(gen[name], "w") as f:
                f.write(text)
        cmd = [os.environ.get("CC", "gcc"), "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-w",
               "-I", os.path.join(HERE, "shim"), "-I", REF, "-I", os.path.join(ROOT, "include"),
               "-o", os.path.join(out_dir, "libimp_ref_gpu.so"),
               gen["filters.c"], os.path.join(REF, "helpers.c"), gen["bridge.c"], gen["advancedio.c"], DROPIN,
               os.path.join(HERE, "ref_stubs.c"), os.path.join(HERE, "fake_freeimage.c"), os.path.join(HERE, "imp_oracle.c"),
               "-L", PKG, "-limp_gpu", "-Wl,-rpath,$ORIGIN/../../ngx_http_imgproc_b200", "-lm"]
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print(os.path.join(out_dir, "libimp_ref_gpu.so"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
