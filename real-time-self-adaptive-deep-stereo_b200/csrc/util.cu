#include "common.cuh"
#include <cstdlib>
#include <mutex>
#include <set>
namespace ms {
static std::mutex g_mu;
static std::string g_err;
void set_error(const std::string& s) { std::lock_guard<std::mutex> l(g_mu); g_err = s; }
const char* last_error_cstr() { std::lock_guard<std::mutex> l(g_mu); return g_err.c_str(); }
static bool g_pdl_suppressed = false;
void pdl_set_suppressed(bool s) { g_pdl_suppressed = s; }
bool pdl_enabled() {
    if (g_pdl_suppressed) return false;
    static int v = -1;
    if (v < 0) { const char* e = getenv("MS_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}
void carveout_once(const void* kernel) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("MS_CARVEOUT"); on = (e && e[0] == '1') ? 1 : 0; }
    if (!on) return;
    static std::set<const void*> seen;
    std::lock_guard<std::mutex> l(g_mu);
    if (!seen.insert(kernel).second) return;
    (void)cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    (void)cudaGetLastError();
}
static long long g_launches = 0;
long long launch_count() { return g_launches; }
void add_launches(long long n) { g_launches += n; }
int check_launch(const char* what, int n_kernels) {
    g_launches += n_kernels;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(std::string(what) + ": " + cudaGetErrorString(e)); return -1; }
    return 0;
}
}  // namespace ms
