/* TEST INFRASTRUCTURE — stand-in for OpenCV 2.4's <opencv/cv.h> (C API), which is not in this
 * image. Declares exactly the types/prototypes the reference names; the
 * definitions are in oracle/ref_stubs.c and forward to the C restatement (oracle/imp_oracle.c)
 * or to cv2 callbacks registered from Python. IplImage keeps the real field order so that
 * sizeof(IplImage)==144 on x86-64 and ctypes can build headers over numpy buffers. */
#ifndef IMP_ORACLE_SHIM_CV_H
#define IMP_ORACLE_SHIM_CV_H
#include <stddef.h>

typedef void CvArr;
typedef struct { int width, height; } CvSize;
typedef struct { int x, y; } CvPoint;
typedef struct { int x, y, width, height; } CvRect;
typedef struct { double val[4]; } CvScalar;
typedef struct { int type; int max_iter; double epsilon; } CvTermCriteria;
typedef struct _IplROI { int coi, xOffset, yOffset, width, height; } IplROI;
typedef struct _IplImage {
    int nSize, ID, nChannels, alphaChannel, depth;
    char colorModel[4], channelSeq[4];
    int dataOrder, origin, align, width, height;
    struct _IplROI* roi;
    struct _IplImage* maskROI;
    void* imageId;
    void* tileInfo;
    int imageSize;
    char* imageData;
    int widthStep;
    int BorderMode[4], BorderConst[4];
    char* imageDataOrigin;
} IplImage;
typedef struct CvMat {
    int type, step;
    int* refcount; int hdr_refcount;
    union { unsigned char* ptr; short* s; int* i; float* fl; double* db; } data;
    int rows, cols;
} CvMat;
typedef struct CvMemStorage CvMemStorage;
typedef struct CvSeq { struct CvSeq* h_next; } CvSeq;
typedef struct CvContour CvContour;

#define IPL_DEPTH_8U  8
#define IPL_DEPTH_32F 32
#define CV_8U     0
#define CV_32SC1  4
#define CV_32FC1  5
#define CV_INTER_NN     0
#define CV_INTER_LINEAR 1
#define CV_INTER_CUBIC  2
#define CV_INTER_AREA   3
#define CV_GAUSSIAN  2
#define CV_BILATERAL 4
#define CV_GRAY2BGR  8
#define CV_IMWRITE_JPEG_QUALITY    1
#define CV_IMWRITE_PNG_COMPRESSION 16
#define CV_TERMCRIT_ITER 1
#define CV_TERMCRIT_EPS  2

static inline CvPoint cvPoint(int x, int y) { CvPoint p; p.x = x; p.y = y; return p; }
static inline CvSize  cvSize(int w, int h) { CvSize s; s.width = w; s.height = h; return s; }
static inline CvRect  cvRect(int x, int y, int w, int h) { CvRect r; r.x = x; r.y = y; r.width = w; r.height = h; return r; }
static inline CvTermCriteria cvTermCriteria(int t, int n, double e) { CvTermCriteria c; c.type = t; c.max_iter = n; c.epsilon = e; return c; }
static inline CvMat cvMat(int rows, int cols, int type, void* data) {
    CvMat m; m.type = type; m.step = cols; m.refcount = 0; m.hdr_refcount = 0;
    m.data.ptr = (unsigned char*)data; m.rows = rows; m.cols = cols; return m;
}

CvSize    cvGetSize(const CvArr* arr);
CvRect    cvGetImageROI(const IplImage* image);
void      cvSetImageROI(IplImage* image, CvRect rect);
IplImage* cvCreateImage(CvSize size, int depth, int channels);
IplImage* cvCreateImageHeader(CvSize size, int depth, int channels);
void      cvReleaseImage(IplImage** image);
void      cvReleaseImageHeader(IplImage** image);
void      cvSetData(CvArr* arr, void* data, int step);
void      cvCopy(const CvArr* src, CvArr* dst, const CvArr* mask);
void      cvResize(const CvArr* src, CvArr* dst, int interpolation);
void      cvFlip(const CvArr* src, CvArr* dst, int flip_mode);
void      cvTranspose(const CvArr* src, CvArr* dst);
void      cvSmooth(const CvArr* src, CvArr* dst, int smoothtype, int p1, int p2, double p3, double p4);
void      cvCvtColor(const CvArr* src, CvArr* dst, int code);
IplImage* cvDecodeImage(const CvMat* buf, int iscolor);
CvMat*    cvEncodeImage(const char* ext, const CvArr* image, const int* params);
CvMat*    cvCreateMat(int rows, int cols, int type);
void      cvReleaseMat(CvMat** mat);
void      cvSetReal2D(CvArr* arr, int i, int j, double v);
double    cvGetReal2D(const CvArr* arr, int i, int j);
int       cvKMeans2(const CvArr* samples, int k, CvArr* labels, CvTermCriteria tc, int attempts,
                    void* rng, int flags, CvArr* centers, double* compactness);
void      cvConvertScale(const CvArr* src, CvArr* dst, double scale, double shift);
#endif
