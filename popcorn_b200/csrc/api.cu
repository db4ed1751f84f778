// Library-level entry points: version, per-thread error text, device query cache.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace pc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED); }

// ---- per-kernel event profiler ------------------------------------------------------------------
constexpr int PROF_MAX_CAT = 64, PROF_MAX_EV = 1 << 16;
static char g_prof_names[PROF_MAX_CAT][64];
static int g_prof_ncat = 0;
static bool g_prof_on = false;
static cudaEvent_t* g_prof_ev = nullptr;     // [2 * PROF_MAX_EV], created lazily
static int g_prof_cat_of[PROF_MAX_EV];
static double g_prof_units[PROF_MAX_EV];
static int g_prof_used = 0, g_prof_created = 0;

int prof_register(const char* name) {
    for (int i = 0; i < g_prof_ncat; ++i)
        if (strncmp(g_prof_names[i], name, 63) == 0) return i;
    if (g_prof_ncat >= PROF_MAX_CAT) return PROF_MAX_CAT - 1;
    strncpy(g_prof_names[g_prof_ncat], name, 63);
    return g_prof_ncat++;
}

ProfScope::ProfScope(int cat, cudaStream_t stream, double units) : slot(-1), st(stream) {
    if (!g_prof_on || g_prof_used >= PROF_MAX_EV) return;
    if (!g_prof_ev) g_prof_ev = (cudaEvent_t*)calloc(2 * PROF_MAX_EV, sizeof(cudaEvent_t));
    slot = g_prof_used++;
    if (slot >= g_prof_created) {
        cudaEventCreate(&g_prof_ev[2 * slot]);
        cudaEventCreate(&g_prof_ev[2 * slot + 1]);
        g_prof_created = slot + 1;
    }
    g_prof_cat_of[slot] = cat;
    g_prof_units[slot] = units;
    cudaEventRecord(g_prof_ev[2 * slot], st);
}
ProfScope::~ProfScope() {
    if (slot >= 0) cudaEventRecord(g_prof_ev[2 * slot + 1], st);
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace pc

extern "C" int pc_version(void) { return 100; }
extern "C" long long pc_launch_count(int reset) {
    const long long v = __atomic_load_n(&pc::g_launches, __ATOMIC_RELAXED);
    if (reset) __atomic_store_n(&pc::g_launches, 0, __ATOMIC_RELAXED);
    return v;
}
extern "C" int pc_profile_enable(int on) {
    pc::g_prof_on = on != 0;
    if (on) pc::g_prof_used = 0;
    return 0;
}
extern "C" int pc_profile_num(void) { return pc::g_prof_ncat; }
extern "C" const char* pc_profile_name(int i) { return (i >= 0 && i < pc::g_prof_ncat) ? pc::g_prof_names[i] : ""; }
extern "C" int pc_profile_get(int cat, double* ms, long long* launches, double* units) {
    PC_CHECK_ARG(ms && launches && units, "null pointer");
    double t = 0, u = 0; long long n = 0;
    for (int s = 0; s < pc::g_prof_used; ++s) {
        if (pc::g_prof_cat_of[s] != cat) continue;
        PC_CUDA(cudaEventSynchronize(pc::g_prof_ev[2 * s + 1]));
        float e = 0.f;
        PC_CUDA(cudaEventElapsedTime(&e, pc::g_prof_ev[2 * s], pc::g_prof_ev[2 * s + 1]));
        t += e; ++n; u += pc::g_prof_units[s];
    }
    *ms = t; *launches = n; *units = u;
    return 0;
}
extern "C" int pc_memcpy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                                 size_t rows, int kind, pc_stream_t stream) {
    PC_CHECK_ARG(dst && src, "null pointer");
    PC_CHECK_ARG(kind == 1 || kind == 2, "kind must be 1 (host->device) or 2 (device->host)");
    if (rows == 0 || width_bytes == 0) return 0;
    PC_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows,
                              kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return 0;
}
extern "C" const char* pc_last_error(void) { return pc::g_err; }

// ---------------------------------------------------------------------------------------------------
// FP32 SIMT ceiling probe (tools/kernel_bench.py): a register-resident FMA loop, scalar or packed x2.
// 2 * 32 * iters FLOP per thread.  Used as the measured "fp32" roofline denominator for the stencils.
// ---------------------------------------------------------------------------------------------------
template <bool X2>
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, float* out) {
    float a[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = threadIdx.x * 1e-3f + i;
    const float m = 1.0000001f, c = 1e-7f;
    if (X2) {
        unsigned long long v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = pc::pack2(a[2 * i], a[2 * i + 1]);
        const unsigned long long mm = pc::pack2(m, m), cc = pc::pack2(c, c);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(mm), "l"(cc));
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) pc::unpack2(v[i], a[2 * i], a[2 * i + 1]);
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = fmaf(a[i], m, c);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += a[i];
    if (s == 12345.678f) out[0] = s;
}

extern "C" int pc_test_fma_peak(int use_x2, int iters, int blocks_per_sm, float* out, pc_stream_t stream) {
    PC_CHECK_ARG(out && iters > 0 && blocks_per_sm > 0, "bad argument");
    const int grid = pc::num_sms() * blocks_per_sm;
    if (use_x2) fma_peak_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(iters, out);
    else fma_peak_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(iters, out);
    PC_LAUNCH_CHECK();
    return grid * 256;
}
