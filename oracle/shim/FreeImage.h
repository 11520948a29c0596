/* TEST INFRASTRUCTURE — stand-in for <FreeImage.h>: only the identifiers bridge.c names
 * (bridge.c:385-474, 642-648, 680-700). The codec itself is out of scope. */
#ifndef IMP_ORACLE_SHIM_FREEIMAGE_H
#define IMP_ORACLE_SHIM_FREEIMAGE_H
#define FREEIMAGE_MAJOR_VERSION 3
#define FREEIMAGE_MINOR_VERSION 18
typedef int FREE_IMAGE_FORMAT;
typedef struct FIMEMORY FIMEMORY;
typedef unsigned char BYTE;
typedef unsigned int DWORD;
enum { FIF_UNKNOWN = -1, FIF_BMP = 0, FIF_JPEG = 2, FIF_TARGA = 17, FIF_TIFF = 18,
       FIF_GIF = 25, FIF_J2K = 30, FIF_JP2 = 31, FIF_WEBP = 35, FIF_JXR = 36 };
#define BMP_SAVE_RLE   1
#define TARGA_SAVE_RLE 2
#define TIFF_DEFLATE   0x0200
#define TIFF_LZW       0x4000
#define TIFF_JPEG      0x8000
#define TIFF_NONE      0x0800
FIMEMORY* FreeImage_OpenMemory(BYTE* data, DWORD size);
void FreeImage_CloseMemory(FIMEMORY* stream);
FREE_IMAGE_FORMAT FreeImage_GetFileTypeFromMemory(FIMEMORY* stream, int size);
FREE_IMAGE_FORMAT FreeImage_GetFIFFromFilename(const char* filename);
#endif
