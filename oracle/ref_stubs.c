/*
 * ref_stubs.c — TEST INFRASTRUCTURE. Link-time stand-ins that let the reference's own
 * filters.c / helpers.c / bridge.c (compiled UNMODIFIED from /root/reference by oracle/Makefile)
 * run in this image, which has neither nginx nor the OpenCV 2.4 C API nor FreeImage.
 *
 *  - IplImage lifetime/ROI functions: minimal re-implementations of the OpenCV C-API contracts
 *    (widthStep = (w*C + 3) & ~3, ROI clipped to the image).
 *  - Pixel-moving OpenCV ops (cvResize, cvSmooth, cvFlip, cvTranspose, cvCopy, cvCvtColor):
 *    forwarded to callbacks registered from Python (cv2 4.13, IPP off — the real OpenCV code path
 *    used for CPU-baseline timing) when set, otherwise to the C restatement in imp_oracle.c.
 *  - Codecs (cvDecodeImage/cvEncodeImage): a RAW container so that RunJob (bridge.c:302-724) can be
 *    driven end to end without a JPEG/PNG codec: blob = 8-byte PNG signature, "IMPR", int32 w,h,c,
 *    then w*h*c tightly packed bytes. Encode writes the same container. Decode/encode are out of
 *    scope for the hot path; only the steps between them are under test.
 *  - FreeImage: oracle/fake_freeimage.c (layout-faithful bitmaps over two trivial containers), so that the
 *    reference's advancedio.c (LoadGIF, LoadSingle, IplToFI32/24) is compiled and driven unmodified too.
 *  - nginx pool functions: malloc/free.
 */
#include "required.h"
#include "advancedio.h"
#include <stdint.h>

typedef struct { unsigned char* data; int width, height, channels, step; } orc_img;
void orc_copy_roi(const orc_img* src, int x, int y, orc_img* dst);
void orc_flip(const orc_img* src, orc_img* dst, int mode);
void orc_transpose(const orc_img* src, orc_img* dst);
void orc_gray2bgr(const orc_img* src, orc_img* dst);
void orc_resize(const orc_img* s, orc_img* d, int mode);
int  orc_gaussian(const orc_img* s, orc_img* d, double sigma);

/* ---- optional cv2 callbacks (set from Python through ctypes) ------------------------------ */
typedef void (*cb_resize_t)(const unsigned char* s, int sw, int sh, int sc, int sstep,
                            unsigned char* d, int dw, int dh, int dstep, int mode);
typedef void (*cb_smooth_t)(unsigned char* s, int w, int h, int c, int step, double sigma);
typedef void (*cb_flip_t)(const unsigned char* s, int w, int h, int c, int sstep, unsigned char* d, int dstep, int mode);
typedef void (*cb_transpose_t)(const unsigned char* s, int w, int h, int c, int sstep, unsigned char* d, int dstep);
static cb_resize_t    g_cb_resize;
static cb_smooth_t    g_cb_smooth;
static cb_flip_t      g_cb_flip;
static cb_transpose_t g_cb_transpose;
void ref_set_callbacks(cb_resize_t r, cb_smooth_t s, cb_flip_t f, cb_transpose_t t) {
    g_cb_resize = r; g_cb_smooth = s; g_cb_flip = f; g_cb_transpose = t;
}

static orc_img view(const IplImage* im) {
    orc_img v;
    v.data = (unsigned char*)im->imageData; v.width = im->width; v.height = im->height;
    v.channels = im->nChannels; v.step = im->widthStep;
    if (im->roi) {
        v.data += (size_t)im->roi->yOffset * im->widthStep + (size_t)im->roi->xOffset * im->nChannels;
        v.width = im->roi->width; v.height = im->roi->height;
    }
    return v;
}

/* ---- IplImage management --------------------------------------------------------------------- */
CvSize cvGetSize(const CvArr* arr) {
    const IplImage* im = (const IplImage*)arr;
    if (im->roi) return cvSize(im->roi->width, im->roi->height);
    return cvSize(im->width, im->height);
}
CvRect cvGetImageROI(const IplImage* im) {
    if (im->roi) return cvRect(im->roi->xOffset, im->roi->yOffset, im->roi->width, im->roi->height);
    return cvRect(0, 0, im->width, im->height);
}
void cvSetImageROI(IplImage* im, CvRect r) {
    int x0 = r.x < 0 ? 0 : r.x, y0 = r.y < 0 ? 0 : r.y;
    int x1 = r.x + r.width > im->width ? im->width : r.x + r.width;
    int y1 = r.y + r.height > im->height ? im->height : r.y + r.height;
    if (!im->roi) im->roi = (IplROI*)calloc(1, sizeof(IplROI));
    im->roi->coi = 0; im->roi->xOffset = x0; im->roi->yOffset = y0;
    im->roi->width = x1 > x0 ? x1 - x0 : 0; im->roi->height = y1 > y0 ? y1 - y0 : 0;
}
IplImage* cvCreateImageHeader(CvSize size, int depth, int channels) {
    IplImage* im = (IplImage*)calloc(1, sizeof(IplImage));
    im->nSize = sizeof(IplImage); im->nChannels = channels; im->depth = depth;
    im->width = size.width; im->height = size.height; im->align = 4;
    im->widthStep = ((size.width * channels * (depth & 255) / 8) + 3) & ~3;
    im->imageSize = im->widthStep * size.height;
    return im;
}
IplImage* cvCreateImage(CvSize size, int depth, int channels) {
    IplImage* im = cvCreateImageHeader(size, depth, channels);
    im->imageData = im->imageDataOrigin = (char*)malloc(im->imageSize > 0 ? (size_t)im->imageSize : 1);
    return im;
}
void cvReleaseImageHeader(IplImage** p) {
    if (p && *p) { free((*p)->roi); free(*p); *p = NULL; }
}
void cvReleaseImage(IplImage** p) {
    if (p && *p) { free((*p)->imageDataOrigin); cvReleaseImageHeader(p); }
}
void cvSetData(CvArr* arr, void* data, int step) {
    IplImage* im = (IplImage*)arr;
    im->imageData = im->imageDataOrigin = (char*)data; im->widthStep = step; im->imageSize = step * im->height;
}

/* ---- pixel-moving ops ------------------------------------------------------------------------ */
void cvCopy(const CvArr* src, CvArr* dst, const CvArr* mask) {
    orc_img s = view((const IplImage*)src), d = view((IplImage*)dst);
    (void)mask;
    orc_copy_roi(&s, 0, 0, &d);
}
void cvResize(const CvArr* src, CvArr* dst, int interpolation) {
    orc_img s = view((const IplImage*)src), d = view((IplImage*)dst);
    if (g_cb_resize) g_cb_resize(s.data, s.width, s.height, s.channels, s.step, d.data, d.width, d.height, d.step, interpolation);
    else orc_resize(&s, &d, interpolation);
}
void cvFlip(const CvArr* src, CvArr* dst, int mode) {
    orc_img s = view((const IplImage*)src), d = view((IplImage*)dst);
    if (g_cb_flip) g_cb_flip(s.data, s.width, s.height, s.channels, s.step, d.data, d.step, mode);
    else orc_flip(&s, &d, mode);
}
void cvTranspose(const CvArr* src, CvArr* dst) {
    orc_img s = view((const IplImage*)src), d = view((IplImage*)dst);
    if (g_cb_transpose) g_cb_transpose(s.data, s.width, s.height, s.channels, s.step, d.data, d.step);
    else orc_transpose(&s, &d);
}
void cvSmooth(const CvArr* src, CvArr* dst, int type, int p1, int p2, double p3, double p4) {
    orc_img s = view((const IplImage*)src), d = view((IplImage*)dst);
    (void)type; (void)p1; (void)p2; (void)p4;
    if (!(p3 > 0)) abort();   /* OpenCV asserts ksize>0 (App. C-9) */
    if (g_cb_smooth && s.data == d.data) g_cb_smooth(s.data, s.width, s.height, s.channels, s.step, p3);
    else orc_gaussian(&s, &d, p3);
}
void cvCvtColor(const CvArr* src, CvArr* dst, int code) {
    orc_img s = view((const IplImage*)src), d = view((IplImage*)dst);
    if (code != CV_GRAY2BGR) abort();
    orc_gray2bgr(&s, &d);
}

/* ---- RAW codec ------------------------------------------------------------------------------- */
IplImage* cvDecodeImage(const CvMat* buf, int iscolor) {
    const unsigned char* b = buf->data.ptr; (void)iscolor;
    if (buf->cols < 24 || memcmp(b + 8, "IMPR", 4) != 0) return NULL;
    int32_t w, h, c; memcpy(&w, b + 12, 4); memcpy(&h, b + 16, 4); memcpy(&c, b + 20, 4);
    if ((long)buf->cols < 24 + (long)w * h * c) return NULL;
    IplImage* im = cvCreateImage(cvSize(w, h), IPL_DEPTH_8U, c);
    for (int y = 0; y < h; y++) memcpy(im->imageData + (size_t)im->widthStep * y, b + 24 + (size_t)y * w * c, (size_t)w * c);
    return im;
}
CvMat* cvEncodeImage(const char* ext, const CvArr* image, const int* params) {
    const IplImage* im = (const IplImage*)image; (void)ext; (void)params;
    int32_t w = im->width, h = im->height, c = im->nChannels;
    size_t n = 24 + (size_t)w * h * c;
    CvMat* m = (CvMat*)calloc(1, sizeof(CvMat));
    m->rows = 1; m->cols = (int)n; m->data.ptr = (unsigned char*)malloc(n);
    memcpy(m->data.ptr, "\x89PNG\r\n\x1a\nIMPR", 12);
    memcpy(m->data.ptr + 12, &w, 4); memcpy(m->data.ptr + 16, &h, 4); memcpy(m->data.ptr + 20, &c, 4);
    for (int y = 0; y < h; y++) memcpy(m->data.ptr + 24 + (size_t)y * w * c, im->imageData + (size_t)im->widthStep * y, (size_t)w * c);
    return m;
}
CvMat* cvCreateMat(int rows, int cols, int type) { (void)rows; (void)cols; (void)type; abort(); }
void cvReleaseMat(CvMat** m) { if (m && *m) { free((*m)->data.ptr); free(*m); *m = NULL; } }
void cvSetReal2D(CvArr* a, int i, int j, double v) { (void)a; (void)i; (void)j; (void)v; abort(); }
double cvGetReal2D(const CvArr* a, int i, int j) { (void)a; (void)i; (void)j; abort(); }
int cvKMeans2(const CvArr* s, int k, CvArr* l, CvTermCriteria tc, int at, void* rng, int fl, CvArr* c, double* cp) {
    (void)s; (void)k; (void)l; (void)tc; (void)at; (void)rng; (void)fl; (void)c; (void)cp; abort();
}
void cvConvertScale(const CvArr* s, CvArr* d, double sc, double sh) { (void)s; (void)d; (void)sc; (void)sh; abort(); }

/* ---- FreeImage: oracle/fake_freeimage.c; advancedio.c itself is the reference's, compiled unmodified ---- */

/* ---- nginx ------------------------------------------------------------------------------------ */
void* ngx_palloc(ngx_pool_t* pool, size_t size) { (void)pool; return malloc(size ? size : 1); }
void* ngx_pnalloc(ngx_pool_t* pool, size_t size) { (void)pool; return malloc(size ? size : 1); }
ngx_int_t ngx_pfree(ngx_pool_t* pool, void* p) { (void)pool; free(p); return 0; }
/* Test URIs are passed already unescaped: plain copy, advancing *dst like nginx does. */
void ngx_unescape_uri(u_char** dst, u_char** src, size_t size, ngx_uint_t type) {
    (void)type; memcpy(*dst, *src, size); *dst += size; *src += size;
}

/* ---- ctypes-friendly entry points ---------------------------------------------------------------- */
/* Runs the reference's RunJob on a RAW blob. Output pixels are copied (tightly packed) into out
 * (capacity out_cap). Returns JobResult.Code; *step = JobResult.Step; dims in *ow,*oh,*oc (0 if none). */
int ref_run_job(const char* uri, const char* exten, const unsigned char* pixels, int w, int h, int c,
                Config* cfg, unsigned char* out, long out_cap, int* ow, int* oh, int* oc, int* step, int* mime) {
    size_t n = 24 + (size_t)w * h * c;
    unsigned char* blob = (unsigned char*)malloc(n);
    int32_t w32 = w, h32 = h, c32 = c;
    memcpy(blob, "\x89PNG\r\n\x1a\nIMPR", 12);
    memcpy(blob + 12, &w32, 4); memcpy(blob + 16, &h32, 4); memcpy(blob + 20, &c32, 4);
    memcpy(blob + 24, pixels, (size_t)w * h * c);
    ngx_connection_t conn; conn.log = NULL;
    ngx_http_request_t req; memset(&req, 0, sizeof req);
    req.connection = &conn;
    req.unparsed_uri.data = (u_char*)uri; req.unparsed_uri.len = strlen(uri);
    req.exten.data = (u_char*)exten; req.exten.len = strlen(exten);
    JobResult* res = RunJob(blob, n, &req, cfg);
    int code = res->Code;
    *step = res->Step; *mime = res->MIME; *ow = *oh = *oc = 0;
    if (code == IMP_OK && res->Length >= 24 && (res->MIME == IMP_MIME_PNG || res->MIME == IMP_MIME_JPG)) {
        int32_t rw, rh, rc; memcpy(&rw, res->EncodedBytes + 12, 4); memcpy(&rh, res->EncodedBytes + 16, 4); memcpy(&rc, res->EncodedBytes + 20, 4);
        *ow = rw; *oh = rh; *oc = rc;
        if ((long)rw * rh * rc <= out_cap) memcpy(out, res->EncodedBytes + 24, (size_t)rw * rh * rc);
        free(res->EncodedBytes);
    }
    free(res); free(blob);
    return code;
}
size_t ref_sizeof_iplimage(void) { return sizeof(IplImage); }
size_t ref_sizeof_config(void) { return sizeof(Config); }

/* ---- advancedio.c entry points -------------------------------------------------------------------- */
/* FiLoadFrames (advancedio.c:311-339) on a container built by the test (IMPGIF1 / IMPFI01, fake_freeimage.c).
 * Returns a heap Album*; *count / *error mirror Album.Count / Album.Error. */
void* ref_fi_load(const unsigned char* buf, int len, int format, int destructive, int page, int* count, int* error) {
    DecodeRequest r;
    r.Buffer = (unsigned char*)buf; r.Length = len; r.Pool = NULL; r.Format = format; r.IsDestructive = destructive; r.Page = page;
    Album* a = (Album*)malloc(sizeof(Album));
    *a = FiLoadFrames(r);
    *count = a->Count; *error = a->Error;
    return a;
}
/* Frame i of an album: geometry + metadata; pixels copied tightly packed into out when out != NULL. */
int ref_album_frame(void* album, int i, int* w, int* h, int* c, int* time, int* dispose, int* key, unsigned char* out) {
    Album* a = (Album*)album;
    if (i < 0 || i >= a->Count || !a->Frames) return -1;
    IplImage* im = a->Frames[i].Image;
    *w = im->width; *h = im->height; *c = im->nChannels;
    *time = a->Frames[i].Time; *dispose = a->Frames[i].Dispose; *key = a->Frames[i].TransparencyKey;
    if (out) for (int y = 0; y < im->height; y++) memcpy(out + (size_t)y * im->width * im->nChannels, im->imageData + (size_t)y * im->widthStep, (size_t)im->width * im->nChannels);
    return 0;
}
void ref_album_free(void* album) {
    Album* a = (Album*)album;
    if (a->Frames) { for (int i = 0; i < a->Count; i++) cvReleaseImage(&a->Frames[i].Image); free(a->Frames); }
    free(a);
}
/* FiSaveFrames (advancedio.c:448-461) of one frame in a non-GIF format: SaveSingle -> IplToFI32 / IplToFI24 -> the fake
 * encoder's dump "IMPFI01\0", int32 w,h,bpp,pitch, then the FIBITMAP's bits as they lie in memory. Returns bytes written. */
long ref_fi_save(const unsigned char* pixels, int w, int h, int c, int step, int format, unsigned char* out, long out_cap, int* error) {
    IplImage* im = cvCreateImageHeader(cvSize(w, h), IPL_DEPTH_8U, c);
    cvSetData(im, (void*)pixels, step);
    Frame fr; fr.Image = im; fr.Time = 0; fr.Dispose = 0; fr.TransparencyKey = 0;
    Album a; a.Frames = &fr; a.Count = 1; a.Error = 0;
    EncodeRequest r; r.Album = &a; r.Pool = NULL; r.Format = format; r.Flags = 0;
    Memory m = FiSaveFrames(r);
    *error = m.Error;
    long n = 0;
    if (!m.Error && m.Length <= out_cap) { memcpy(out, m.Buffer, (size_t)m.Length); n = m.Length; }
    if (!m.Error) free(m.Buffer);
    cvReleaseImageHeader(&im);
    return n;
}
/* RunJob on an arbitrary blob (e.g. the fake GIF container, so that decode goes through the reference's FiLoadFrames and
 * encode through cvEncodeImage's RAW container or FiSaveFrames' dump). The encoded bytes are copied to out. */
int ref_run_job_blob(const char* uri, const char* exten, const unsigned char* blob, long len, Config* cfg,
                     unsigned char* out, long out_cap, long* out_len, int* step, int* mime) {
    unsigned char* copy = (unsigned char*)malloc((size_t)len);
    memcpy(copy, blob, (size_t)len);
    ngx_connection_t conn; conn.log = NULL;
    ngx_http_request_t req; memset(&req, 0, sizeof req);
    req.connection = &conn;
    req.unparsed_uri.data = (u_char*)uri; req.unparsed_uri.len = strlen(uri);
    req.exten.data = (u_char*)exten; req.exten.len = strlen(exten);
    JobResult* res = RunJob(copy, (size_t)len, &req, cfg);
    int code = res->Code;
    *step = res->Step; *mime = res->MIME; *out_len = 0;
    if (code == IMP_OK && res->EncodedBytes && (long)res->Length <= out_cap) {
        memcpy(out, res->EncodedBytes, res->Length);
        *out_len = (long)res->Length;
        free(res->EncodedBytes);
    }
    free(res); free(copy);
    return code;
}
