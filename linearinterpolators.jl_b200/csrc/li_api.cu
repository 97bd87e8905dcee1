// li_api.cu -- library plumbing: errors, device, memory, vscale!, launch counter.
#include <stdarg.h>
#include <stdlib.h>

#include "li_common.cuh"

namespace li {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
std::atomic<int> g_options[OPT_COUNT];

static const char *const kOptionNames[OPT_COUNT] = {"no_fused", "no_stream_dn", "no_bands", "affine_plain",
                                                    "trace", "adj_debug", "no_multi", "exp"};
static const char *const kOptionEnv[OPT_COUNT] = {"LI_B200_NO_FUSED", "LI_B200_NO_STREAM_DN", "LI_B200_NO_BANDS",
                                                  "LI_B200_AFFINE_PLAIN", "LI_B200_TRACE", "LI_B200_ADJ_DEBUG",
                                                  "LI_B200_NO_MULTI", "LI_B200_EXP"};
// the environment is consulted once, at load time
static const bool g_options_init = [] {
    for (int k = 0; k < OPT_COUNT; ++k) {
        const char *v = getenv(kOptionEnv[k]);
        g_options[k].store(v ? (atoi(v) ? atoi(v) : 1) : 0);
    }
    return true;
}();

static int option_index(const char *name)
{
    if (name)
        for (int k = 0; k < OPT_COUNT; ++k)
            if (!strcmp(name, kOptionNames[k])) return k;
    return -1;
}

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

// vscale!(y, beta): beta == 1 no-op, beta == 0 zero-fill (y is not read), else scale.
template <typename T>
__global__ void vscale_kernel(T *__restrict__ y, int64_t n, T beta)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] *= beta;
}

int vscale_impl(int dtype, void *y, int64_t n, double beta, cudaStream_t st)
{
    if (n <= 0 || beta == 1.0) return LI_OK;
    if (beta == 0.0) {
        LI_CUDA(cudaMemsetAsync(y, 0, (size_t)n * dtype_size(dtype), st));
        return LI_OK;
    }
    const int threads = 256;
    const unsigned blocks = (unsigned)std::min<int64_t>((n + threads - 1) / threads, 148 * 16);
    if (dtype == LI_F32)
        vscale_kernel<float><<<blocks, threads, 0, st>>>((float *)y, n, (float)beta);
    else
        vscale_kernel<double><<<blocks, threads, 0, st>>>((double *)y, n, beta);
    LI_CHECK_LAUNCH();
    return LI_OK;
}

// Grow-only device staging buffers for the host-buffer entry points, one set per (host thread, stream):
// cudaMalloc/cudaFree per call would dominate small applies and serialise the device, and calls on different
// streams must not share buffers.
struct StageSet {
    // 0: src, 1: dst, 2: workspace, 3: partial accumulators, 4: frame pointers, 5: partial models of frame groups,
    // 6: rank-private accumulator of the all-reduce, 7: vector-verb partial sums, 8: positions of an axis-aligned affine
    // operator, 9: scalars of regularize!
    void *ptr[10] = {};
    size_t bytes[10] = {};
};
static thread_local std::map<std::pair<int, void *>, StageSet> g_stage;   // key: (device, stream)

int staging(int slot, size_t bytes, void **out, void *stream_key)
{
    int dev = 0;
    LI_CUDA(cudaGetDevice(&dev));
    StageSet &s = g_stage[std::make_pair(dev, stream_key)];
    if (bytes > s.bytes[slot]) {
        if (s.ptr[slot]) {
            cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream_key));   // nothing in flight may still use it
            cudaFree(s.ptr[slot]);
        }
        s.ptr[slot] = nullptr;
        s.bytes[slot] = 0;
        const size_t want = bytes + bytes / 8 + 256;
        LI_CUDA(cudaMalloc(&s.ptr[slot], want));
        s.bytes[slot] = want;
    }
    *out = s.ptr[slot];
    return LI_OK;
}

}  // namespace li

extern "C" {

int li_version(void) { return LI_B200_VERSION; }

const char *li_last_error(void) { return li::g_err; }

const char *li_status_string(int status)
{
    switch (status) {
    case LI_OK: return "LI_OK";
    case LI_ERR_ARG: return "LI_ERR_ARG";
    case LI_ERR_DIM: return "LI_ERR_DIM";
    case LI_ERR_INEXACT: return "LI_ERR_INEXACT";
    case LI_ERR_BOUNDS: return "LI_ERR_BOUNDS";
    case LI_ERR_UNSUPPORTED: return "LI_ERR_UNSUPPORTED";
    case LI_ERR_CUDA: return "LI_ERR_CUDA";
    case LI_ERR_NOMEM: return "LI_ERR_NOMEM";
    case LI_ERR_NCCL: return "LI_ERR_NCCL";
    default: return "LI_ERR_UNKNOWN";
    }
}

int li_device_count(int *count)
{
    LI_REQUIRE(count, LI_ERR_ARG, "null argument");
    LI_CUDA(cudaGetDeviceCount(count));
    return LI_OK;
}

int li_set_device(int device)
{
    LI_CUDA(cudaSetDevice(device));
    return LI_OK;
}

int li_get_device(int *device)
{
    LI_REQUIRE(device, LI_ERR_ARG, "null argument");
    LI_CUDA(cudaGetDevice(device));
    return LI_OK;
}

int li_num_sms(int *count)
{
    LI_REQUIRE(count, LI_ERR_ARG, "null argument");
    int dev = 0;
    LI_CUDA(cudaGetDevice(&dev));
    LI_CUDA(cudaDeviceGetAttribute(count, cudaDevAttrMultiProcessorCount, dev));
    return LI_OK;
}

int li_malloc(void **dev_ptr, size_t bytes)
{
    LI_REQUIRE(dev_ptr, LI_ERR_ARG, "null argument");
    LI_CUDA(cudaMalloc(dev_ptr, bytes ? bytes : 1));
    return LI_OK;
}

int li_free(void *dev_ptr)
{
    LI_CUDA(cudaFree(dev_ptr));
    return LI_OK;
}

int li_host_alloc(void **host_ptr, size_t bytes)
{
    LI_REQUIRE(host_ptr, LI_ERR_ARG, "null argument");
    LI_CUDA(cudaHostAlloc(host_ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return LI_OK;
}

int li_host_free(void *host_ptr)
{
    LI_CUDA(cudaFreeHost(host_ptr));
    return LI_OK;
}

int li_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes, void *stream)
{
    LI_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, li::as_stream(stream)));
    return LI_OK;
}

int li_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes, void *stream)
{
    LI_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, li::as_stream(stream)));
    return LI_OK;
}

int li_memcpy_d2d(void *dst_dev, const void *src_dev, size_t bytes, void *stream)
{
    LI_CUDA(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, li::as_stream(stream)));
    return LI_OK;
}

int li_memset_zero(void *dst_dev, size_t bytes, void *stream)
{
    LI_CUDA(cudaMemsetAsync(dst_dev, 0, bytes, li::as_stream(stream)));
    return LI_OK;
}

int li_stream_create(void **stream)
{
    LI_REQUIRE(stream, LI_ERR_ARG, "null argument");
    cudaStream_t st;
    LI_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *stream = st;
    return LI_OK;
}

int li_stream_destroy(void *stream)
{
    LI_CUDA(cudaStreamDestroy(li::as_stream(stream)));
    return LI_OK;
}

int li_stream_sync(void *stream)
{
    LI_CUDA(cudaStreamSynchronize(li::as_stream(stream)));
    return LI_OK;
}

int li_vscale(int dtype, void *y_dev, int64_t n, double beta, void *stream)
{
    LI_REQUIRE(dtype == LI_F32 || dtype == LI_F64, LI_ERR_ARG, "bad dtype");
    LI_REQUIRE(y_dev || n == 0, LI_ERR_ARG, "null argument");
    return li::vscale_impl(dtype, y_dev, n, beta, li::as_stream(stream));
}

int64_t li_kernel_launch_count(void) { return (int64_t)li::g_launches.load(); }

int li_set_option(const char *name, int value)
{
    const int k = li::option_index(name);
    LI_REQUIRE(k >= 0, LI_ERR_ARG, "unknown option '%s'", name ? name : "(null)");
    li::g_options[k].store(value);
    return LI_OK;
}

int li_get_option(const char *name, int *value)
{
    const int k = li::option_index(name);
    LI_REQUIRE(k >= 0 && value, LI_ERR_ARG, "unknown option '%s'", name ? name : "(null)");
    *value = li::g_options[k].load();
    return LI_OK;
}

}  // extern "C"
