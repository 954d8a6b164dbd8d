// placeholder until the vocoder lands (keeps the ABI complete)
#include "engine.cuh"
using namespace jv;
struct jv_hift { Engine eng; };
extern "C" {
int jv_hift_create(int, int, jv_hift**) { set_last_error("hift not built yet"); return JV_ERR_STATE; }
void jv_hift_destroy(jv_hift*) {}
int jv_hift_set_weight(jv_hift*, const char*, const float*, const int64_t*, int) { return JV_ERR_STATE; }
int jv_hift_finalize(jv_hift*) { return JV_ERR_STATE; }
size_t jv_hift_workspace_bytes(const jv_hift*, int, const int32_t*) { return 0; }
int jv_hift_f0(jv_hift*, int, int, const int32_t*, const float*, float*, void*, size_t, void*) { return JV_ERR_STATE; }
int jv_hift_source(jv_hift*, int, int, const int32_t*, const float*, const float*, const float*, float*, void*) { return JV_ERR_STATE; }
int jv_hift_decode(jv_hift*, int, int, const int32_t*, const float*, const float*, float*, void*, size_t, void*) { return JV_ERR_STATE; }
}
