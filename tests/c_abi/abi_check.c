/* Compiled by tests/test_abi.py against include/b2lz4.h with a C compiler (the header's prototypes are checked by gcc,
 * not only looked up by name through ctypes) and linked with libb2lz4.so.  It is the C twin of zig/lz4.zig: every
 * entry point the Zig shim binds is called here with the argument types the shim uses.
 * With a CUDA device it round-trips a block, an HC block and a frame; without one every compute call must fail loudly
 * with B2LZ4_ERR_CUDA (there is no CPU fallback).  Prints "abi ok gpu" / "abi ok nogpu". */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b2lz4.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "abi_check failed at line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main(void) {
    const size_t n = 200000;
    uint8_t* src = (uint8_t*)malloc(n);
    for (size_t i = 0; i < n; i++) src[i] = (uint8_t)((i * 7) % 61 + (i / 4096));
    const size_t bound = b2lz4_compress_bound(n);
    CHECK(bound == n + n / 255 + 16);
    CHECK(b2lz4_compress_bound((size_t)0x7E000001u) == 0);
    uint8_t* comp = (uint8_t*)malloc(bound);
    uint8_t* back = (uint8_t*)malloc(n);
    size_t out = 0, out2 = 0;
    CHECK(strcmp(b2lz4_status_name(B2LZ4_OK), "ok") == 0);
    CHECK(b2lz4_version() != NULL);
    CHECK(b2lz4_debug_tune("no_such_knob", 1) == -1);

    b2lz4f_prefs prefs;
    b2lz4f_prefs_init(&prefs);
    prefs.block_mode = 1; prefs.block_checksum = 1; prefs.content_checksum = 1; prefs.content_size = n;
    const size_t fbound = b2lz4f_compress_frame_bound(n, &prefs);
    CHECK(fbound > n);
    uint8_t* frame = (uint8_t*)malloc(fbound);
    uint8_t hdr[19];   /* lz4f.HEADER_SIZE_MAX, src/lz4f.zig:20 */
    size_t hsize = 0;
    CHECK(b2lz4f_write_frame_header(hdr, sizeof hdr, &prefs, &hsize) == B2LZ4_OK && hsize == 15);
    b2lz4f_prefs info; size_t psize = 0;
    CHECK(b2lz4f_parse_frame_header(hdr, hsize, &info, &psize) == B2LZ4_OK && psize == hsize && info.content_size == n);
    CHECK(b2lz4f_header_size(hdr, hsize, &psize) == B2LZ4_OK && psize == 15);

    int rc = b2lz4_compress_default(src, n, comp, bound, &out);
    if (rc == B2LZ4_ERR_CUDA) {
        /* no device: everything that computes must say so */
        CHECK(b2lz4_decompress_safe(comp, 10, back, n, &out2) == B2LZ4_ERR_CUDA);
        CHECK(b2lz4_compress_hc(src, n, comp, bound, 9, &out) == B2LZ4_ERR_CUDA);
        CHECK(b2lz4f_compress_frame(src, n, frame, fbound, &prefs, &out) == B2LZ4_ERR_CUDA);
        CHECK(b2lz4_last_cuda_error() != NULL);
        printf("abi ok nogpu\n");
        return 0;
    }
    CHECK(rc == B2LZ4_OK && out > 0 && out < n);
    CHECK(b2lz4_decompress_safe(comp, out, back, n, &out2) == B2LZ4_OK && out2 == n && memcmp(src, back, n) == 0);
    CHECK(b2lz4_decompress_safe(comp, out, back, n - 1, &out2) == B2LZ4_ERR_OUTPUT_TOO_SMALL);
    CHECK(b2lz4_compress_fast(src, n, comp, bound, 7, &out) == B2LZ4_OK);
    CHECK(b2lz4_compress_hc(src, n, comp, bound, 9, &out) == B2LZ4_OK && out < n);
    memset(back, 0, n);
    CHECK(b2lz4_decompress_safe(comp, out, back, n, &out2) == B2LZ4_OK && out2 == n && memcmp(src, back, n) == 0);
    CHECK(b2lz4_compress_hc(src, n, comp, bound, 12, &out) == B2LZ4_ERR_UNSUPPORTED_LEVEL);
    size_t consumed = n;
    CHECK(b2lz4_compress_dest_size(src, comp, 1000, &consumed, &out) == B2LZ4_OK && out <= 1000 && consumed < n);
    uint32_t h = 0;
    CHECK(b2lz4_xxh32("", 0, 0, &h) == B2LZ4_OK && h == 0x02CC5D05u);
    CHECK(b2lz4f_compress_frame(src, n, frame, fbound, &prefs, &out) == B2LZ4_OK && out > hsize + 8);
    memset(back, 0, n);
    CHECK(b2lz4f_decompress_frame(frame, out, back, n, &out2) == B2LZ4_OK && out2 == n && memcmp(src, back, n) == 0);
    frame[out - 1] ^= 1;
    CHECK(b2lz4f_decompress_frame(frame, out, back, n, &out2) == B2LZ4F_ERR_CONTENT_CHECKSUM_INVALID);
    /* the README streaming trio */
    b2lz4f_cctx* cctx = NULL;
    CHECK(b2lz4f_create_compression_context(&cctx) == B2LZ4_OK && cctx != NULL);
    size_t pos = 0, w = 0;
    CHECK(b2lz4f_compress_begin(cctx, frame, fbound, &prefs, &w) == B2LZ4_OK); pos += w;
    CHECK(b2lz4f_compress_update(cctx, frame + pos, fbound - pos, src, n / 2, &w) == B2LZ4_OK); pos += w;
    CHECK(b2lz4f_compress_update(cctx, frame + pos, fbound - pos, src + n / 2, n - n / 2, &w) == B2LZ4_OK); pos += w;
    CHECK(b2lz4f_compress_end(cctx, frame + pos, fbound - pos, &w) == B2LZ4_OK); pos += w;
    b2lz4f_free_compression_context(cctx);
    memset(back, 0, n);
    CHECK(b2lz4f_decompress_frame(frame, pos, back, n, &out2) == B2LZ4_OK && out2 == n && memcmp(src, back, n) == 0);
    CHECK(b2lz4_kernel_launch_count() > 0);
    printf("abi ok gpu\n");
    return 0;
}
