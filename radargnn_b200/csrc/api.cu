// api.cu -- library-level entry points of the C ABI (include/rgnn.h): version, status
// text, device probe, launch counter.
#include <atomic>
#include <mutex>
#include <string>
#include <vector>
#include <string.h>

#include "common.cuh"

namespace rgnn {
namespace {
thread_local char g_last_error[512] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace

void set_last_cuda_error(cudaError_t err, const char* file, int line) {
  snprintf(g_last_error, sizeof(g_last_error), "%s (%s) at %s:%d", cudaGetErrorName(err),
           cudaGetErrorString(err), file, line);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- per-kernel profiling ---------------------------------------------------------------
namespace {
struct ProfileRecord { const char* name; cudaEvent_t start, stop; };
std::atomic<int> g_profile_on{0};
std::mutex g_profile_mu;
std::vector<ProfileRecord> g_profile_records;
struct ProfileTotal { std::string name; double ms; int64_t count; };
std::vector<ProfileTotal> g_profile_totals;
}  // namespace

ProfileScope::ProfileScope(const char* name, cudaStream_t stream)
    : name_(name), stream_(stream), start_(nullptr), active_(g_profile_on.load(std::memory_order_relaxed) != 0) {
  if (!active_) return;
  if (cudaEventCreate(&start_) != cudaSuccess || cudaEventRecord(start_, stream_) != cudaSuccess) active_ = false;
}

ProfileScope::~ProfileScope() {
  if (!active_) return;
  cudaEvent_t stop;
  if (cudaEventCreate(&stop) != cudaSuccess) return;
  cudaEventRecord(stop, stream_);
  std::lock_guard<std::mutex> lock(g_profile_mu);
  g_profile_records.push_back({name_, start_, stop});
}

}  // namespace rgnn

extern "C" {

int rgnn_abi_version(void) { return RGNN_ABI_VERSION; }

const char* rgnn_status_string(int status) {
  switch (status) {
    case RGNN_OK: return "ok";
    case RGNN_ERR_INVALID_ARGUMENT: return "invalid argument";
    case RGNN_ERR_K_NOT_SMALLER_THAN_N: return "Expected n_neighbors < n_samples_fit";
    case RGNN_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    case RGNN_ERR_CUDA: return "CUDA error";
    case RGNN_ERR_DOT_PRODUCT: return "Error in dot product calculation";
    case RGNN_ERR_INVALID_FEATURE: return "Invalid feature specified";
    case RGNN_ERR_UNSUPPORTED: return "unsupported configuration";
    case RGNN_ERR_NO_DEVICE: return "no CUDA device";
    case RGNN_ERR_NON_FINITE_INPUT: return "Input contains NaN or infinity";
    case RGNN_ERR_INDEX_OUT_OF_RANGE: return "edge_index holds a node id outside [0, N)";
    default: return "unknown status";
  }
}

const char* rgnn_last_cuda_error(void) { return rgnn::g_last_error; }

int rgnn_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return RGNN_ERR_NO_DEVICE;
  }
  int dev = 0;
  RGNN_CUDA_CHECK(cudaGetDevice(&dev));
  int sms = 0, major = 0, minor = 0;
  RGNN_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  RGNN_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  RGNN_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = major;
  if (cc_minor) *cc_minor = minor;
  return RGNN_OK;
}

int64_t rgnn_kernel_launch_count(void) { return rgnn::g_launches.load(std::memory_order_relaxed); }

void rgnn_profile_enable(int32_t on) {
  rgnn::g_profile_on.store(on ? 1 : 0, std::memory_order_relaxed);
}

int32_t rgnn_profile_collect(void) {
  using namespace rgnn;
  std::lock_guard<std::mutex> lock(g_profile_mu);
  for (const ProfileRecord& r : g_profile_records) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.stop) == cudaSuccess && cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
      bool found = false;
      for (ProfileTotal& t : g_profile_totals)
        if (t.name == r.name) { t.ms += ms; t.count += 1; found = true; break; }
      if (!found) g_profile_totals.push_back({r.name, ms, 1});
    }
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  g_profile_records.clear();
  cudaGetLastError();
  return static_cast<int32_t>(g_profile_totals.size());
}

int rgnn_profile_entry(int32_t index, const char** name, double* total_ms, int64_t* count) {
  using namespace rgnn;
  std::lock_guard<std::mutex> lock(g_profile_mu);
  if (index < 0 || index >= static_cast<int32_t>(g_profile_totals.size())) return RGNN_ERR_INVALID_ARGUMENT;
  if (name) *name = g_profile_totals[index].name.c_str();
  if (total_ms) *total_ms = g_profile_totals[index].ms;
  if (count) *count = g_profile_totals[index].count;
  return RGNN_OK;
}

void rgnn_profile_reset(void) {
  using namespace rgnn;
  rgnn_profile_collect();
  std::lock_guard<std::mutex> lock(g_profile_mu);
  g_profile_totals.clear();
}

}  // extern "C"
