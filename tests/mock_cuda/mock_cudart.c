/* mock_cudart.c -- TEST INFRASTRUCTURE: a host-memory stand-in for the CUDA runtime symbols that
 * libtorchain_b200's objects import, so that the CPU suite can drive the library's HOST logic (handles,
 * per-device uploads, workspaces, launch sequence, the chunk loops of the host-buffer entry) end to end on a
 * machine without a GPU.  "Device" memory is calloc'ed host memory, copies are memcpy, streams and events are
 * tokens, and kernel launches are only recorded (name, grid, block, first 1 KiB of the first parameter) --
 * nothing is computed.  Never linked into the product: tests/mock_cuda/Makefile links the SAME object files
 * as libtorchain_b200.so against this file instead of libcudart (tests/test_mock_host.py). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef struct { unsigned x, y, z; } dim3;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };

/* ---- bookkeeping the tests read back ---- */
#define MAX_FUNCS 512
#define MAX_LAUNCHES 4096
#define PARAM_BYTES 2048
static struct { const void* host_fun; const char* name; } g_funcs[MAX_FUNCS];
static int g_nfuncs = 0;
typedef struct {
  const char* name;
  unsigned grid, block;
  int cooperative;
  void* stream;
  size_t smem;
  unsigned char param0[PARAM_BYTES];
} MockLaunch;
static MockLaunch g_launches[MAX_LAUNCHES];
static int g_nlaunches = 0, g_nlaunches_total = 0;
static __thread int g_last_slot = 0;
static long g_mallocs = 0, g_frees = 0, g_live_bytes = 0, g_memcpys = 0, g_host_allocs = 0;
static __thread int g_device = 0; /* the current device is per host thread, as in CUDA */
static int g_device_count = 1;
static long g_streams_live = 0;

/* every allocation carries its size in a 256-byte header so that live bytes can be tracked */
static int g_fill = 0; /* byte every new allocation is filled with (device memory is not zeroed in real life either) */
void mock_set_fill(int byte) { g_fill = byte; }
static void* tracked_alloc(size_t n) {
  unsigned char* p = (unsigned char*)calloc(1, n + 256);
  if (!p) return NULL;
  *(size_t*)p = n;
  if (g_fill) memset(p + 256, g_fill, n);
  return p + 256;
}
static size_t tracked_free(void* q) {
  unsigned char* p = (unsigned char*)q - 256;
  size_t n = *(size_t*)p;
  free(p);
  return n;
}

int mock_launch_count(void) { return g_nlaunches; }
void mock_reset_launches(void) { g_nlaunches = 0; g_nlaunches_total = 0; }
const char* mock_launch_name(int i) { return (i >= 0 && i < g_nlaunches && g_launches[i].name) ? g_launches[i].name : ""; }
int mock_launch_info(int i, unsigned* grid, unsigned* block, int* cooperative, void** stream, unsigned long long* smem) {
  if (i < 0 || i >= g_nlaunches) return 0;
  if (grid) *grid = g_launches[i].grid;
  if (block) *block = g_launches[i].block;
  if (cooperative) *cooperative = g_launches[i].cooperative;
  if (stream) *stream = g_launches[i].stream;
  if (smem) *smem = g_launches[i].smem;
  return 1;
}
const void* mock_launch_param0(int i) { return (i >= 0 && i < g_nlaunches) ? g_launches[i].param0 : NULL; }
long mock_stat(int what) {
  switch (what) {
    case 0: return g_mallocs;
    case 1: return g_frees;
    case 2: return g_live_bytes;
    case 3: return g_memcpys;
    case 4: return g_host_allocs;
    case 5: return g_streams_live;
    default: return -1;
  }
}
void mock_set_device_count(int n) { g_device_count = n; }

/* ---- registration (called from the objects' static constructors) ---- */
void** __cudaRegisterFatBinary(void* fatCubin) { (void)fatCubin; static void* handle; return &handle; }
void __cudaRegisterFatBinaryEnd(void** h) { (void)h; }
void __cudaUnregisterFatBinary(void** h) { (void)h; }
void __cudaRegisterFunction(void** h, const char* hostFun, char* deviceFun, const char* deviceName, int thread_limit,
                            void* tid, void* bid, void* bDim, void* gDim, int* wSize) {
  (void)h; (void)deviceFun; (void)thread_limit; (void)tid; (void)bid; (void)bDim; (void)gDim; (void)wSize;
  if (g_nfuncs < MAX_FUNCS) {
    g_funcs[g_nfuncs].host_fun = hostFun;
    g_funcs[g_nfuncs].name = deviceName;
    g_nfuncs++;
  }
}
static __thread dim3 g_cfg_grid, g_cfg_block;
static __thread size_t g_cfg_smem;
static __thread void* g_cfg_stream;
unsigned __cudaPushCallConfiguration(dim3 grid, dim3 block, size_t smem, void* stream) {
  g_cfg_grid = grid; g_cfg_block = block; g_cfg_smem = smem; g_cfg_stream = stream;
  return 0;
}
cudaError_t __cudaPopCallConfiguration(dim3* grid, dim3* block, size_t* smem, void* stream) {
  *grid = g_cfg_grid; *block = g_cfg_block; *smem = g_cfg_smem; *(void**)stream = g_cfg_stream;
  return cudaSuccess;
}
static cudaError_t record_launch(const void* func, dim3 grid, dim3 block, void** args, size_t smem, void* stream, int coop) {
  /* launches may come from several host threads: reserve a slot atomically; when the log is full the
   * last slot is overwritten (the count keeps running, readers clamp) */
  int slot = __atomic_fetch_add(&g_nlaunches_total, 1, __ATOMIC_RELAXED);
  if (slot >= MAX_LAUNCHES) slot = MAX_LAUNCHES - 1;
  else __atomic_fetch_add(&g_nlaunches, 1, __ATOMIC_RELAXED);
  MockLaunch* l = &g_launches[slot];
  g_last_slot = slot;
  l->name = "?";
  for (int i = 0; i < g_nfuncs; i++)
    if (g_funcs[i].host_fun == func) l->name = g_funcs[i].name;
  l->grid = grid.x; l->block = block.x; l->cooperative = coop; l->stream = stream; l->smem = smem;
  memset(l->param0, 0, PARAM_BYTES);
  (void)args;
  return cudaSuccess;
}
/* The size of a kernel's parameter is not known here; the library's parameter structs are copied by the
 * test-side helper mock_capture_param (below) through a size the test supplies. */
static size_t g_param_bytes = 0;
void mock_set_param_capture(size_t nbytes) { g_param_bytes = nbytes > PARAM_BYTES ? PARAM_BYTES : nbytes; }
/* per-kernel override: kernels whose name contains `pattern` capture `nbytes` (their parameter struct's size) */
#define MAX_RULES 8
static struct { char pattern[64]; size_t nbytes; } g_rules[MAX_RULES];
static int g_nrules = 0;
void mock_set_param_capture_for(const char* pattern, size_t nbytes) {
  if (g_nrules >= MAX_RULES) return;
  strncpy(g_rules[g_nrules].pattern, pattern, 63);
  g_rules[g_nrules].nbytes = nbytes > PARAM_BYTES ? PARAM_BYTES : nbytes;
  g_nrules++;
}
static size_t capture_bytes(const char* name) {
  for (int i = 0; i < g_nrules; i++)
    if (name && strstr(name, g_rules[i].pattern)) return g_rules[i].nbytes;
  return g_param_bytes;
}
cudaError_t cudaLaunchKernel(const void* func, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t stream) {
  cudaError_t e = record_launch(func, grid, block, args, smem, stream, 0);
  if (args && args[0] && capture_bytes(g_launches[g_last_slot].name)) memcpy(g_launches[g_last_slot].param0, args[0], capture_bytes(g_launches[g_last_slot].name));
  return e;
}
cudaError_t cudaLaunchCooperativeKernel(const void* func, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t stream) {
  cudaError_t e = record_launch(func, grid, block, args, smem, stream, 1);
  if (args && args[0] && capture_bytes(g_launches[g_last_slot].name)) memcpy(g_launches[g_last_slot].param0, args[0], capture_bytes(g_launches[g_last_slot].name));
  return e;
}
cudaError_t cudaFuncSetAttribute(const void* f, int attr, int value) { (void)f; (void)attr; (void)value; return cudaSuccess; }

/* ---- devices ---- */
cudaError_t cudaGetDevice(int* d) { *d = g_device; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= g_device_count) return cudaErrorInvalidValue; g_device = d; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = g_device_count; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, int attr, int dev) { (void)attr; (void)dev; *v = 148; return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "mock CUDA error"; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }

/* ---- memory ---- */
cudaError_t cudaMalloc(void** p, size_t n) {
  *p = tracked_alloc(n ? n : 1);
  if (!*p) return cudaErrorMemoryAllocation;
  __atomic_fetch_add(&g_mallocs, 1, __ATOMIC_RELAXED); __atomic_fetch_add(&g_live_bytes, (long)n, __ATOMIC_RELAXED);
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) { if (p) { __atomic_fetch_sub(&g_live_bytes, (long)tracked_free(p), __ATOMIC_RELAXED); __atomic_fetch_add(&g_frees, 1, __ATOMIC_RELAXED); } return cudaSuccess; }
/* The host-buffer entry polls a mapped pinned counter that the backward kernel advances.  With
 * mock_set_auto_progress(limit) every cudaStreamQuery stands in for one frame of kernel progress: it bumps the
 * most recently allocated MAPPED pinned word and answers "not ready" until the counter has reached `limit`. */
static __thread volatile unsigned* g_mapped_word = NULL; /* the calling thread's own counter */
static unsigned g_auto_progress = 0;
void mock_set_auto_progress(unsigned limit) { g_auto_progress = limit; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned flags) {
  *p = tracked_alloc(n ? n : 1);
  __atomic_fetch_add(&g_host_allocs, 1, __ATOMIC_RELAXED);
  if (*p && (flags & 2u) && n == sizeof(unsigned)) g_mapped_word = (volatile unsigned*)*p;
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFreeHost(void* p) {
  if (p == (void*)g_mapped_word) g_mapped_word = NULL;
  if (p) tracked_free(p);
  return cudaSuccess;
}
cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned flags) { (void)flags; *d = h; return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int kind) { (void)kind; memmove(d, s, n); __atomic_fetch_add(&g_memcpys, 1, __ATOMIC_RELAXED); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int kind, cudaStream_t st) { (void)st; return cudaMemcpy(d, s, n, kind); }
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, int kind, cudaStream_t st) {
  (void)kind; (void)st;
  for (size_t r = 0; r < h; r++) memmove((char*)d + r * dp, (const char*)s + r * sp, w);
  g_memcpys++;
  return cudaSuccess;
}
cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st) { (void)st; memset(d, v, n); return cudaSuccess; }

/* ---- streams and events: tokens; everything "completes" at once ---- */
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags) { (void)flags; *s = malloc(8); __atomic_fetch_add(&g_streams_live, 1, __ATOMIC_RELAXED); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); __atomic_fetch_sub(&g_streams_live, 1, __ATOMIC_RELAXED); return cudaSuccess; }
cudaError_t cudaStreamQuery(cudaStream_t s) {
  (void)s;
  if (g_auto_progress && g_mapped_word && *g_mapped_word < g_auto_progress) {
    (*g_mapped_word)++;
    return 600; /* cudaErrorNotReady */
  }
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t s) { (void)s; return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags) { (void)s; (void)e; (void)flags; return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = malloc(8); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags) { (void)flags; *e = malloc(8); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) { (void)e; (void)s; return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t e) { (void)e; return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t e) { (void)e; return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { (void)a; (void)b; *ms = 1.0f; return cudaSuccess; }

/* ---- driver entry point: a cuStreamWriteValue32 that performs the write at once ---- */
static int mock_write_value32(void* stream, unsigned long long addr, unsigned value, unsigned flags) {
  (void)stream; (void)flags;
  *(volatile unsigned*)(uintptr_t)addr = value;
  return 0;
}
cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** fn, unsigned long long flags, int* status) {
  (void)flags;
  *fn = strcmp(symbol, "cuStreamWriteValue32") == 0 ? (void*)mock_write_value32 : NULL;
  if (status) *status = *fn ? 0 : 1;
  return cudaSuccess;
}
