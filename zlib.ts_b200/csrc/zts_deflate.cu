// placeholder: replaced by the real pipeline
#include "zts_common.cuh"
extern "C" {
int zlb_deflate_batch(zlb_ctx* ctx, const void*, void*, const zlb_item*, zlb_result*, size_t, int, int, uint32_t, uint32_t)
{ return zts_fail(ctx, ZLB_E_UNSUPPORTED, "deflate not built yet"); }
int zlb_deflate_batch_host(zlb_ctx* ctx, const void*, size_t, void*, size_t, const zlb_item*, zlb_result*, size_t, int, int, uint32_t, uint32_t)
{ return zts_fail(ctx, ZLB_E_UNSUPPORTED, "deflate not built yet"); }
uint64_t zlb_deflate_bound(uint64_t n, uint32_t, int) { return n * 2 + 1024; }
int zlb_debug_lz77(zlb_ctx* ctx, const void*, uint32_t, uint32_t*, uint32_t*, uint32_t*) { return zts_fail(ctx, ZLB_E_UNSUPPORTED, "n/a"); }
int zlb_debug_code_lengths(zlb_ctx* ctx, const uint32_t*, int, int, uint8_t*) { return zts_fail(ctx, ZLB_E_UNSUPPORTED, "n/a"); }
}
