/* TEST INFRASTRUCTURE — stand-in for nginx's <ngx_config.h>, just enough for the
 * reference's required.h:12-14 to resolve so that /root/reference/{filters,helpers,bridge}.c
 * compile UNMODIFIED into oracle/_ref/libimp_ref.so (see oracle/Makefile).
 * Nothing here is nginx code; only the names the reference touches are declared. */
#ifndef IMP_ORACLE_SHIM_NGX_CONFIG_H
#define IMP_ORACLE_SHIM_NGX_CONFIG_H
#include <stdlib.h>
#include <string.h>
#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>

typedef intptr_t  ngx_int_t;
typedef uintptr_t ngx_uint_t;
typedef intptr_t  ngx_flag_t;
typedef struct { size_t len; u_char* data; } ngx_str_t;
typedef struct ngx_pool_s ngx_pool_t;
typedef struct { void* log; } ngx_connection_t;
typedef struct {
    ngx_connection_t* connection;
    ngx_pool_t*       pool;
    ngx_str_t         unparsed_uri;
    ngx_str_t         exten;
} ngx_http_request_t;

#define NGX_LOG_ERR 4
#define ngx_log_error(level, log, err, ...) ((void)0)

void* ngx_palloc(ngx_pool_t* pool, size_t size);
void* ngx_pnalloc(ngx_pool_t* pool, size_t size);
ngx_int_t ngx_pfree(ngx_pool_t* pool, void* p);
void  ngx_unescape_uri(u_char** dst, u_char** src, size_t size, ngx_uint_t type);
#endif
