// wunet_tc.cu — bf16 / tcgen05 path (placeholder until the tensor-core kernels land).
#include "wunet_tc.cuh"
#include <cstdio>

namespace wunet {
struct TcState { int dummy; };
static thread_local char g_tc_err[256] = "";
const char *tc_error() { return g_tc_err; }
int tc_set_weights(TcState **st, int, int, const TcBlockSrc *, int, const float *, const float *, cudaStream_t)
{
    if (!*st) *st = new TcState{0};
    return 0;
}
size_t tc_workspace_bytes(int, int, int, int) { return 256; }
int tc_forward(TcState *, const float *, float *, int, int, void *, cudaStream_t, int *, cudaEvent_t *)
{
    snprintf(g_tc_err, sizeof(g_tc_err), "bf16 tcgen05 path not built yet");
    return -1;
}
int tc_read_level(TcState *, int, const void *, int, int, float *, cudaStream_t)
{
    snprintf(g_tc_err, sizeof(g_tc_err), "bf16 tcgen05 path not built yet");
    return -1;
}
void tc_destroy(TcState *st) { delete st; }
}  // namespace wunet
