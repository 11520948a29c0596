/* TEST INFRASTRUCTURE — stand-in for <FreeImage.h>: the identifiers bridge.c (bridge.c:385-474, 642-648,
 * 680-700) and advancedio.c name, so that both reference files compile unmodified. The functions are
 * implemented by oracle/fake_freeimage.c over two trivial in-memory container formats (no real codec:
 * decode/encode are out of scope); what is kept from the real library is its memory
 * layout — bottom-up scanlines, rows padded to 4 bytes, BGRA byte order, palette as RGBQUAD. */
#ifndef IMP_ORACLE_SHIM_FREEIMAGE_H
#define IMP_ORACLE_SHIM_FREEIMAGE_H
#define FREEIMAGE_MAJOR_VERSION 3
#define FREEIMAGE_MINOR_VERSION 18
typedef int FREE_IMAGE_FORMAT;
typedef int FREE_IMAGE_COLOR_TYPE;
typedef int FREE_IMAGE_MDMODEL;
typedef int FREE_IMAGE_MDTYPE;
typedef int FREE_IMAGE_QUANTIZE;
typedef int BOOL;
typedef struct FIMEMORY FIMEMORY;
typedef struct FIBITMAP FIBITMAP;
typedef struct FIMULTIBITMAP FIMULTIBITMAP;
typedef struct FITAG FITAG;
typedef unsigned char BYTE;
typedef unsigned short WORD;
typedef unsigned int DWORD;
typedef struct { BYTE rgbBlue, rgbGreen, rgbRed, rgbReserved; } RGBQUAD;
enum { FIF_UNKNOWN = -1, FIF_BMP = 0, FIF_ICO = 1, FIF_JPEG = 2, FIF_JNG = 3, FIF_KOALA = 4, FIF_LBM = 5, FIF_IFF = 5,
       FIF_MNG = 6, FIF_PBM = 7, FIF_PBMRAW = 8, FIF_PCD = 9, FIF_PCX = 10, FIF_PGM = 11, FIF_PGMRAW = 12,
       FIF_PNG = 13, FIF_PPM = 14, FIF_PPMRAW = 15, FIF_RAS = 16, FIF_TARGA = 17, FIF_TIFF = 18, FIF_WBMP = 19,
       FIF_PSD = 20, FIF_CUT = 21, FIF_XBM = 22, FIF_XPM = 23, FIF_DDS = 24, FIF_GIF = 25, FIF_HDR = 26,
       FIF_FAXG3 = 27, FIF_SGI = 28, FIF_EXR = 29, FIF_J2K = 30, FIF_JP2 = 31, FIF_PFM = 32, FIF_PICT = 33,
       FIF_RAW = 34, FIF_WEBP = 35, FIF_JXR = 36 };
enum { FIC_MINISWHITE = 0, FIC_MINISBLACK = 1, FIC_RGB = 2, FIC_PALETTE = 3, FIC_RGBALPHA = 4, FIC_CMYK = 5 };
enum { FIMD_COMMENTS = 0, FIMD_ANIMATION = 9 };
enum { FIDT_BYTE = 1, FIDT_SHORT = 3, FIDT_LONG = 4 };
enum { FIQ_WUQUANT = 0, FIQ_NNQUANT = 1 };
#define BMP_SAVE_RLE   1
#define TARGA_SAVE_RLE 2
#define TIFF_DEFLATE   0x0200
#define TIFF_LZW       0x4000
#define TIFF_JPEG      0x8000
#define TIFF_NONE      0x0800
FIMEMORY* FreeImage_OpenMemory(BYTE* data, DWORD size);
void FreeImage_CloseMemory(FIMEMORY* stream);
BOOL FreeImage_AcquireMemory(FIMEMORY* stream, BYTE** data, DWORD* size);
FREE_IMAGE_FORMAT FreeImage_GetFileTypeFromMemory(FIMEMORY* stream, int size);
FREE_IMAGE_FORMAT FreeImage_GetFIFFromFilename(const char* filename);
FIBITMAP* FreeImage_Allocate(int width, int height, int bpp, unsigned rmask, unsigned gmask, unsigned bmask);
void FreeImage_Unload(FIBITMAP* dib);
FIBITMAP* FreeImage_LoadFromMemory(FREE_IMAGE_FORMAT fif, FIMEMORY* stream, int flags);
BOOL FreeImage_SaveToMemory(FREE_IMAGE_FORMAT fif, FIBITMAP* dib, FIMEMORY* stream, int flags);
FIMULTIBITMAP* FreeImage_LoadMultiBitmapFromMemory(FREE_IMAGE_FORMAT fif, FIMEMORY* stream, int flags);
BOOL FreeImage_SaveMultiBitmapToMemory(FREE_IMAGE_FORMAT fif, FIMULTIBITMAP* bitmap, FIMEMORY* stream, int flags);
BOOL FreeImage_CloseMultiBitmap(FIMULTIBITMAP* bitmap, int flags);
int FreeImage_GetPageCount(FIMULTIBITMAP* bitmap);
void FreeImage_AppendPage(FIMULTIBITMAP* bitmap, FIBITMAP* data);
FIBITMAP* FreeImage_LockPage(FIMULTIBITMAP* bitmap, int page);
void FreeImage_UnlockPage(FIMULTIBITMAP* bitmap, FIBITMAP* data, BOOL changed);
unsigned FreeImage_GetWidth(FIBITMAP* dib);
unsigned FreeImage_GetHeight(FIBITMAP* dib);
unsigned FreeImage_GetBPP(FIBITMAP* dib);
unsigned FreeImage_GetPitch(FIBITMAP* dib);
BYTE* FreeImage_GetBits(FIBITMAP* dib);
BYTE* FreeImage_GetScanLine(FIBITMAP* dib, int scanline);
RGBQUAD* FreeImage_GetPalette(FIBITMAP* dib);
FREE_IMAGE_COLOR_TYPE FreeImage_GetColorType(FIBITMAP* dib);
int FreeImage_GetTransparentIndex(FIBITMAP* dib);
void FreeImage_SetTransparentIndex(FIBITMAP* dib, int index);
void FreeImage_SetTransparent(FIBITMAP* dib, BOOL enabled);
FIBITMAP* FreeImage_ConvertTo8Bits(FIBITMAP* dib);
FIBITMAP* FreeImage_ConvertTo24Bits(FIBITMAP* dib);
FIBITMAP* FreeImage_ConvertTo32Bits(FIBITMAP* dib);
FIBITMAP* FreeImage_ColorQuantizeEx(FIBITMAP* dib, FREE_IMAGE_QUANTIZE quantize, int PaletteSize, int ReserveSize, RGBQUAD* ReservePalette);
BOOL FreeImage_GetMetadata(FREE_IMAGE_MDMODEL model, FIBITMAP* dib, const char* key, FITAG** tag);
BOOL FreeImage_SetMetadata(FREE_IMAGE_MDMODEL model, FIBITMAP* dib, const char* key, FITAG* tag);
FITAG* FreeImage_CreateTag(void);
void FreeImage_DeleteTag(FITAG* tag);
const char* FreeImage_GetTagKey(FITAG* tag);
const void* FreeImage_GetTagValue(FITAG* tag);
BOOL FreeImage_SetTagKey(FITAG* tag, const char* key);
BOOL FreeImage_SetTagType(FITAG* tag, FREE_IMAGE_MDTYPE type);
BOOL FreeImage_SetTagCount(FITAG* tag, DWORD count);
BOOL FreeImage_SetTagLength(FITAG* tag, DWORD length);
BOOL FreeImage_SetTagValue(FITAG* tag, const void* value);
#endif
