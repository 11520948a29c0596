/* TEST INFRASTRUCTURE — a stand-in FreeImage for the compiled reference (oracle/_ref/libimp_ref.so).
 *
 * The reference's advancedio.c (LoadGIF advancedio.c:104-262, LoadSingle :264-309, IplToFI32/24 :65-101,
 * SaveGIF/SaveSingle :341-446) is compiled UNMODIFIED against oracle/shim/FreeImage.h; this file provides the
 * library underneath it. There is no real codec here (decode/encode are out of scope): "files"
 * are two trivial containers the tests write and read directly,
 *
 *   IMPGIF1\0  int32 pages | per page: int32 w,h,left,top,dispose,key,time | 256 x {B,G,R,x} | w*h indices, top-down
 *   IMPFI01\0  int32 w,h,bpp,pitch | pitch*h bytes exactly as FreeImage holds them (bottom-up scanlines)
 *
 * What IS kept from the real library is the in-memory layout the reference's loops depend on: scanline 0 is the
 * bottom row, rows are padded to 4 bytes and zero-filled on allocation, 24/32-bit pixels are B,G,R[,A], the
 * palette is an RGBQUAD[256] that directly follows a BITMAPINFOHEADER whose last field (biClrImportant) is 256 for
 * 8-bit bitmaps — so `palette[-1]`, which LoadGIF reads for a page without a transparent colour, is the bytes
 * {0,1,0,0} (FreeImage 3.x BitmapAccess.cpp layout, restated from its published source; not verifiable here). */
#include <stdlib.h>
#include <string.h>
#include "FreeImage.h"

struct FITAG { char key[32]; int type; unsigned count, length; unsigned char value[16]; };

struct FIBITMAP {
    int w, h, bpp, pitch;
    BYTE* bits;
    unsigned clr_important;          /* directly before pal[]: what palette[-1] aliases */
    RGBQUAD pal[256];
    int key;                         /* transparent index, -1 = none */
    int has[4];                      /* FrameTime, DisposalMethod, FrameLeft, FrameTop present */
    struct FITAG tag[4];
};

struct FIMEMORY { BYTE* data; DWORD size, cap; int owns; };

struct FIMULTIBITMAP { int count, cap; FIBITMAP** pages; };

static const char* const TAG_KEYS[4] = {"FrameTime", "DisposalMethod", "FrameLeft", "FrameTop"};
static const char GIF_MAGIC[8] = "IMPGIF1";
static const char FI_MAGIC[8] = "IMPFI01";

static int rd32(const BYTE* p) { int v; memcpy(&v, p, 4); return v; }

static void mem_write(FIMEMORY* m, const void* p, size_t n) {
    if (m->size + n > m->cap) {
        m->cap = (DWORD)((m->size + n) * 2 + 64);
        m->data = (BYTE*)realloc(m->data, m->cap);
    }
    memcpy(m->data + m->size, p, n);
    m->size += (DWORD)n;
}
static void mem_w32(FIMEMORY* m, int v) { mem_write(m, &v, 4); }

/* ---- memory streams ---- */
FIMEMORY* FreeImage_OpenMemory(BYTE* data, DWORD size) {
    FIMEMORY* m = (FIMEMORY*)calloc(1, sizeof(*m));
    if (data) { m->data = data; m->size = size; m->cap = size; m->owns = 0; }
    else m->owns = 1;
    return m;
}
void FreeImage_CloseMemory(FIMEMORY* m) { if (!m) return; if (m->owns) free(m->data); free(m); }
BOOL FreeImage_AcquireMemory(FIMEMORY* m, BYTE** data, DWORD* size) { *data = m->data; *size = m->size; return 1; }
FREE_IMAGE_FORMAT FreeImage_GetFileTypeFromMemory(FIMEMORY* m, int n) {
    (void)n;
    if (m && m->size >= 8 && !memcmp(m->data, GIF_MAGIC, 8)) return FIF_GIF;
    if (m && m->size >= 8 && !memcmp(m->data, FI_MAGIC, 8)) return FIF_BMP;
    return FIF_UNKNOWN;
}
FREE_IMAGE_FORMAT FreeImage_GetFIFFromFilename(const char* f) {
    if (!f) return FIF_UNKNOWN;
    const char* dot = strrchr(f, '.');
    const char* e = dot ? dot + 1 : f;
    if (!strcmp(e, "gif")) return FIF_GIF;
    if (!strcmp(e, "bmp")) return FIF_BMP;
    if (!strcmp(e, "tga")) return FIF_TARGA;
    if (!strcmp(e, "tif") || !strcmp(e, "tiff")) return FIF_TIFF;
    if (!strcmp(e, "webp")) return FIF_WEBP;
    if (!strcmp(e, "jp2")) return FIF_JP2;
    if (!strcmp(e, "ppm")) return FIF_PPM;
    return FIF_UNKNOWN;
}

/* ---- bitmaps ---- */
FIBITMAP* FreeImage_Allocate(int w, int h, int bpp, unsigned r, unsigned g, unsigned b) {
    (void)r; (void)g; (void)b;
    FIBITMAP* d = (FIBITMAP*)calloc(1, sizeof(*d));
    d->w = w; d->h = h; d->bpp = bpp;
    d->pitch = ((w * bpp + 7) / 8 + 3) & ~3;
    d->bits = (BYTE*)calloc((size_t)d->pitch * (h > 0 ? h : 1) + 1, 1);
    d->clr_important = bpp <= 8 ? (1u << bpp) : 0;
    d->key = -1;
    return d;
}
static FIBITMAP* clone(FIBITMAP* s) {
    FIBITMAP* d = FreeImage_Allocate(s->w, s->h, s->bpp, 0, 0, 0);
    memcpy(d->bits, s->bits, (size_t)s->pitch * s->h);
    memcpy(d->pal, s->pal, sizeof(d->pal));
    d->key = s->key;
    memcpy(d->has, s->has, sizeof(d->has));
    memcpy(d->tag, s->tag, sizeof(d->tag));
    return d;
}
void FreeImage_Unload(FIBITMAP* d) { if (!d) return; free(d->bits); free(d); }
unsigned FreeImage_GetWidth(FIBITMAP* d) { return (unsigned)d->w; }
unsigned FreeImage_GetHeight(FIBITMAP* d) { return (unsigned)d->h; }
unsigned FreeImage_GetBPP(FIBITMAP* d) { return (unsigned)d->bpp; }
unsigned FreeImage_GetPitch(FIBITMAP* d) { return (unsigned)d->pitch; }
BYTE* FreeImage_GetBits(FIBITMAP* d) { return d->bits; }
BYTE* FreeImage_GetScanLine(FIBITMAP* d, int y) { return d->bits + (size_t)d->pitch * y; }
RGBQUAD* FreeImage_GetPalette(FIBITMAP* d) { return d->bpp <= 8 ? d->pal : NULL; }
FREE_IMAGE_COLOR_TYPE FreeImage_GetColorType(FIBITMAP* d) { return d->bpp == 32 ? FIC_RGBALPHA : d->bpp == 24 ? FIC_RGB : FIC_PALETTE; }
int FreeImage_GetTransparentIndex(FIBITMAP* d) { return d->key; }
void FreeImage_SetTransparentIndex(FIBITMAP* d, int i) { d->key = i; }
void FreeImage_SetTransparent(FIBITMAP* d, BOOL on) { if (!on) d->key = -1; }

FIBITMAP* FreeImage_ConvertTo8Bits(FIBITMAP* s) { return clone(s); }      /* pages of the fake container are always 8-bit */
static FIBITMAP* convert(FIBITMAP* s, int bpp) {
    FIBITMAP* d = FreeImage_Allocate(s->w, s->h, bpp, 0, 0, 0);
    const int nb = bpp / 8;
    for (int y = 0; y < s->h; y++) {
        const BYTE* sp = s->bits + (size_t)s->pitch * y;
        BYTE* dp = d->bits + (size_t)d->pitch * y;
        for (int x = 0; x < s->w; x++) {
            BYTE px[4] = {0, 0, 0, 255};
            if (s->bpp == 8) { RGBQUAD q = s->pal[sp[x]]; px[0] = q.rgbBlue; px[1] = q.rgbGreen; px[2] = q.rgbRed; px[3] = (sp[x] == s->key) ? 0 : 255; }
            else memcpy(px, sp + x * (s->bpp / 8), (size_t)(s->bpp / 8));
            memcpy(dp + x * nb, px, (size_t)nb);
        }
    }
    return d;
}
FIBITMAP* FreeImage_ConvertTo24Bits(FIBITMAP* s) { return convert(s, 24); }
FIBITMAP* FreeImage_ConvertTo32Bits(FIBITMAP* s) { return convert(s, 32); }
/* no quantiser: a 3-3-2 bit palette, enough for SaveGIF to run end to end (its output is never compared) */
FIBITMAP* FreeImage_ColorQuantizeEx(FIBITMAP* s, FREE_IMAGE_QUANTIZE q, int n, int rn, RGBQUAD* rp) {
    (void)q; (void)n; (void)rn; (void)rp;
    FIBITMAP* d = FreeImage_Allocate(s->w, s->h, 8, 0, 0, 0);
    for (int i = 0; i < 256; i++) { d->pal[i].rgbRed = (BYTE)((i >> 5) * 36); d->pal[i].rgbGreen = (BYTE)(((i >> 2) & 7) * 36); d->pal[i].rgbBlue = (BYTE)((i & 3) * 85); }
    for (int y = 0; y < s->h; y++)
        for (int x = 0; x < s->w; x++) {
            const BYTE* p = s->bits + (size_t)s->pitch * y + x * (s->bpp / 8);
            int idx = ((p[2] >> 5) << 5) | ((p[1] >> 5) << 2) | (p[0] >> 6);
            d->bits[(size_t)d->pitch * y + x] = (BYTE)(idx == 255 ? 254 : idx);
        }
    return d;
}

/* ---- metadata (animation model only) ---- */
static int tag_slot(const char* key) { for (int i = 0; i < 4; i++) if (!strcmp(key, TAG_KEYS[i])) return i; return -1; }
BOOL FreeImage_GetMetadata(FREE_IMAGE_MDMODEL model, FIBITMAP* d, const char* key, FITAG** tag) {
    int s = tag_slot(key);
    if (model != FIMD_ANIMATION || s < 0 || !d->has[s]) { *tag = NULL; return 0; }
    *tag = &d->tag[s];
    return 1;
}
BOOL FreeImage_SetMetadata(FREE_IMAGE_MDMODEL model, FIBITMAP* d, const char* key, FITAG* tag) {
    int s = tag_slot(key);
    if (model != FIMD_ANIMATION || s < 0) return 0;
    d->tag[s] = *tag; d->has[s] = 1;
    return 1;
}
FITAG* FreeImage_CreateTag(void) { return (FITAG*)calloc(1, sizeof(FITAG)); }
void FreeImage_DeleteTag(FITAG* t) { free(t); }
const char* FreeImage_GetTagKey(FITAG* t) { return t->key; }
const void* FreeImage_GetTagValue(FITAG* t) { return t->value; }
BOOL FreeImage_SetTagKey(FITAG* t, const char* k) { strncpy(t->key, k, sizeof(t->key) - 1); return 1; }
BOOL FreeImage_SetTagType(FITAG* t, FREE_IMAGE_MDTYPE ty) { t->type = ty; return 1; }
BOOL FreeImage_SetTagCount(FITAG* t, DWORD c) { t->count = c; return 1; }
BOOL FreeImage_SetTagLength(FITAG* t, DWORD l) { t->length = l; return 1; }
BOOL FreeImage_SetTagValue(FITAG* t, const void* v) { memset(t->value, 0, sizeof(t->value)); memcpy(t->value, v, t->length < 16 ? t->length : 16); return 1; }
static void set_tag(FIBITMAP* d, int slot, const void* v, unsigned len) {
    struct FITAG* t = &d->tag[slot];
    memset(t, 0, sizeof(*t));
    strcpy(t->key, TAG_KEYS[slot]); t->length = len; t->count = 1;
    memcpy(t->value, v, len);
    d->has[slot] = 1;
}

/* ---- single bitmaps: IMPFI01 ---- */
FIBITMAP* FreeImage_LoadFromMemory(FREE_IMAGE_FORMAT fif, FIMEMORY* m, int flags) {
    (void)fif; (void)flags;
    if (!m || m->size < 24 || memcmp(m->data, FI_MAGIC, 8)) return NULL;
    int w = rd32(m->data + 8), h = rd32(m->data + 12), bpp = rd32(m->data + 16), pitch = rd32(m->data + 20);
    FIBITMAP* d = FreeImage_Allocate(w, h, bpp, 0, 0, 0);
    if (pitch != d->pitch || m->size < 24 + (DWORD)(pitch * h)) { FreeImage_Unload(d); return NULL; }
    memcpy(d->bits, m->data + 24, (size_t)pitch * h);
    return d;
}
BOOL FreeImage_SaveToMemory(FREE_IMAGE_FORMAT fif, FIBITMAP* d, FIMEMORY* m, int flags) {
    (void)fif; (void)flags;
    mem_write(m, FI_MAGIC, 8);
    mem_w32(m, d->w); mem_w32(m, d->h); mem_w32(m, d->bpp); mem_w32(m, d->pitch);
    mem_write(m, d->bits, (size_t)d->pitch * d->h);
    return 1;
}

/* ---- multi-page: IMPGIF1 ---- */
FIMULTIBITMAP* FreeImage_LoadMultiBitmapFromMemory(FREE_IMAGE_FORMAT fif, FIMEMORY* m, int flags) {
    (void)fif; (void)flags;
    FIMULTIBITMAP* c = (FIMULTIBITMAP*)calloc(1, sizeof(*c));
    if (!m || m->size == 0) return c;                              /* SaveGIF opens an empty stream to append to */
    if (m->size < 12 || memcmp(m->data, GIF_MAGIC, 8)) { free(c); return NULL; }
    const int n = rd32(m->data + 8);
    c->pages = (FIBITMAP**)calloc((size_t)(n > 0 ? n : 1), sizeof(FIBITMAP*));
    c->cap = n;
    const BYTE* p = m->data + 12;
    for (int i = 0; i < n; i++) {
        int w = rd32(p), h = rd32(p + 4), left = rd32(p + 8), top = rd32(p + 12), dispose = rd32(p + 16), key = rd32(p + 20), time = rd32(p + 24);
        p += 28;
        FIBITMAP* d = FreeImage_Allocate(w, h, 8, 0, 0, 0);
        memcpy(d->pal, p, 1024); p += 1024;
        for (int y = 0; y < h; y++) memcpy(d->bits + (size_t)d->pitch * (h - 1 - y), p + (size_t)y * w, (size_t)w);
        p += (size_t)w * h;
        d->key = key;
        long t = time; short l = (short)left, tp = (short)top; int dv = dispose;
        set_tag(d, 0, &t, sizeof(long)); set_tag(d, 1, &dv, sizeof(int)); set_tag(d, 2, &l, sizeof(short)); set_tag(d, 3, &tp, sizeof(short));
        c->pages[c->count++] = d;
    }
    return c;
}
BOOL FreeImage_CloseMultiBitmap(FIMULTIBITMAP* c, int flags) {
    (void)flags;
    if (!c) return 0;
    for (int i = 0; i < c->count; i++) FreeImage_Unload(c->pages[i]);
    free(c->pages); free(c);
    return 1;
}
int FreeImage_GetPageCount(FIMULTIBITMAP* c) { return c->count; }
FIBITMAP* FreeImage_LockPage(FIMULTIBITMAP* c, int page) { return (page >= 0 && page < c->count) ? c->pages[page] : NULL; }
void FreeImage_UnlockPage(FIMULTIBITMAP* c, FIBITMAP* d, BOOL changed) { (void)c; (void)d; (void)changed; }
void FreeImage_AppendPage(FIMULTIBITMAP* c, FIBITMAP* d) {
    if (c->count == c->cap) { c->cap = c->cap ? c->cap * 2 : 8; c->pages = (FIBITMAP**)realloc(c->pages, (size_t)c->cap * sizeof(FIBITMAP*)); }
    c->pages[c->count++] = clone(d);
}
BOOL FreeImage_SaveMultiBitmapToMemory(FREE_IMAGE_FORMAT fif, FIMULTIBITMAP* c, FIMEMORY* m, int flags) {
    (void)fif; (void)flags;
    mem_write(m, GIF_MAGIC, 8);
    mem_w32(m, c->count);
    for (int i = 0; i < c->count; i++) {
        FIBITMAP* d = c->pages[i];
        int time = 0, dispose = 0;
        if (d->has[0]) memcpy(&time, d->tag[0].value, 4);
        if (d->has[1]) dispose = d->tag[1].value[0];
        mem_w32(m, d->w); mem_w32(m, d->h); mem_w32(m, 0); mem_w32(m, 0); mem_w32(m, dispose); mem_w32(m, d->key); mem_w32(m, time);
        mem_write(m, d->pal, 1024);
        for (int y = 0; y < d->h; y++) mem_write(m, d->bits + (size_t)d->pitch * (d->h - 1 - y), (size_t)d->w);
    }
    return 1;
}
