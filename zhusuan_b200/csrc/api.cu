// C-ABI plumbing: version, thread-local error string, launch checking.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = {0};

void zsb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int zsb_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    zsb_set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return ZSB_ERR_CUDA;
  }
  return ZSB_OK;
}

extern "C" {

int zsb_version(void) { return 100; }  // 0.1.0

// Copies the calling thread's last error message (NUL-terminated) into buf; returns its length.
int zsb_last_error(char* buf, size_t n) {
  if (buf && n) {
    strncpy(buf, g_err, n - 1);
    buf[n - 1] = 0;
  }
  return (int)strlen(g_err);
}

// Number of CUDA devices visible (0 when there is no GPU: every compute entry point then fails
// with ZSB_ERR_CUDA -- there is no CPU fallback in this library).
int zsb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// Blocks until `stream` drains; surfaces asynchronous kernel faults as an error code.
int zsb_stream_sync(void* stream) {
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e != cudaSuccess) {
    zsb_set_error("zsb_stream_sync: %s", cudaGetErrorString(e));
    return ZSB_ERR_CUDA;
  }
  return ZSB_OK;
}

}  // extern "C"
