// C ABI of libucdir_b200.so: op dispatcher, error reporting, capability probe.  See include/ucdir_b200.h.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <nvtx3/nvToolsExt.h>
#include "common.cuh"

namespace ucdir {
static thread_local char g_err[512] = "";
long long g_launches = 0;
// per-op CUDA-event profiling (bench.py's live roofline measurement): one event after every op
static bool g_prof = false;
static std::vector<cudaEvent_t> g_ev;
static std::vector<int> g_ev_op;      // index of the op (within its run_ops call) the event closes, -1 = call start
static std::vector<int> g_ev_kind;
static size_t g_ev_used = 0;
static void prof_mark(cudaStream_t st, int op_index, int kind) {
  if (g_ev_used == g_ev.size()) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return; g_ev.push_back(e); g_ev_op.push_back(0); g_ev_kind.push_back(0); }
  g_ev_op[g_ev_used] = op_index; g_ev_kind[g_ev_used] = kind;
  cudaEventRecord(g_ev[g_ev_used++], st);
}
int sm_count() {
  static int n_dev[UCDIR_MAX_DEV] = {};
  const int d = cur_dev();
  if (!n_dev[d]) { cudaDeviceGetAttribute(&n_dev[d], cudaDevAttrMultiProcessorCount, d); if (n_dev[d] <= 0) n_dev[d] = 148; }
  return n_dev[d];
}
void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}

// NVTX ranges (SURVEY 5 "tracing"): with UCDIR_NVTX=1 every op issued by ucdir_run_ops / captured by ucdir_graph_capture is
// wrapped in a range named after its kind and shape, and graph replays in "ucdir step graph" -- visible in nsys / ncu --nvtx.
static const bool g_nvtx = []() { const char* e = getenv("UCDIR_NVTX"); return e && e[0] == '1'; }();
static const char* kind_name(int k) {
  switch (k) {
    case UCDIR_OP_CONV_F32: return "conv_f32"; case UCDIR_OP_SGEMM_F32: return "sgemm_f32"; case UCDIR_OP_SOFTMAX_F32: return "softmax";
    case UCDIR_OP_GUIDANCE: return "guidance"; case UCDIR_OP_TIME_EMBED: return "time_embed"; case UCDIR_OP_GATHER_TILES: return "gather_tiles";
    case UCDIR_OP_SCATTER: return "scatter+posterior"; case UCDIR_OP_MAXPOOL2: return "maxpool2"; case UCDIR_OP_MEMSET: return "memset";
    case UCDIR_OP_TC_CONV: return "tc_conv"; case UCDIR_OP_TC_ATTN: return "tc_attn"; case UCDIR_OP_GN_APPLY_BF16: return "gn_apply";
    case UCDIR_OP_CAST: return "cast"; case UCDIR_OP_CROP_TILES: return "crop_tiles"; case UCDIR_OP_GN_STATS_F32: return "gn_stats";
    case UCDIR_OP_GN_APPLY_F32: return "gn_apply_f32"; case UCDIR_OP_LAYOUT: return "layout"; case UCDIR_OP_TO_IMAGE_U8: return "to_image_u8";
    default: return "op";
  }
}
struct NvtxOp {
  bool on;
  NvtxOp(const ucdir_op_t& op, int index) : on(g_nvtx) {
    if (!on) return;
    char name[160];
    if (op.kind == UCDIR_OP_TC_CONV)
      snprintf(name, sizeof(name), "#%d tc_conv %dx%dx%d C%d+%d->%d taps%d g%d mode%d%s", index, op.i[UCDIR_TC_I_B], op.i[UCDIR_TC_I_H], op.i[UCDIR_TC_I_W],
               op.i[UCDIR_TC_I_C0], op.i[UCDIR_TC_I_C1], op.i[UCDIR_TC_I_NTOT], op.i[UCDIR_TC_I_NTY] * op.i[UCDIR_TC_I_NTX], op.i[UCDIR_TC_I_GROUPS],
               op.i[UCDIR_TC_I_MODE], op.i[UCDIR_TC_I_SPLIT] ? " split" : "");
    else if (op.kind == UCDIR_OP_CONV_F32)
      snprintf(name, sizeof(name), "#%d conv_f32 %dx%dx%d C%d+%d->%d k%d g%d mode%d", index, op.i[UCDIR_CONV_I_B], op.i[UCDIR_CONV_I_H], op.i[UCDIR_CONV_I_W],
               op.i[UCDIR_CONV_I_C0], op.i[UCDIR_CONV_I_C1], op.i[UCDIR_CONV_I_COUT], op.i[UCDIR_CONV_I_KSIZE], op.i[UCDIR_CONV_I_GROUPS], op.i[UCDIR_CONV_I_MODE]);
    else if (op.kind == UCDIR_OP_TC_ATTN)
      snprintf(name, sizeof(name), "#%d tc_attn B%d N%d C%d", index, op.i[UCDIR_ATTN_I_B], op.i[UCDIR_ATTN_I_N], op.i[UCDIR_ATTN_I_C]);
    else snprintf(name, sizeof(name), "#%d %s", index, kind_name(op.kind));
    nvtxRangePushA(name);
  }
  ~NvtxOp() { if (on) nvtxRangePop(); }
};

static int dispatch(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  switch (op.kind) {
    case UCDIR_OP_CONV_F32: return launch_conv_f32(op, st, dry);
    case UCDIR_OP_SGEMM_F32: return launch_sgemm_f32(op, st, dry);
    case UCDIR_OP_SOFTMAX_F32: return launch_softmax_f32(op, st, dry);
    case UCDIR_OP_GUIDANCE: return launch_guidance(op, st, dry);
    case UCDIR_OP_TIME_EMBED: return launch_time_embed(op, st, dry);
    case UCDIR_OP_GATHER_TILES: return launch_gather_tiles(op, st, dry);
    case UCDIR_OP_SCATTER: return launch_scatter(op, st, dry);
    case UCDIR_OP_MAXPOOL2: return launch_maxpool2(op, st, dry);
    case UCDIR_OP_MEMSET: {
      size_t n = (size_t)op.i[0] + ((size_t)op.i[1] << 31);
      if (!op.p[0]) { set_error("memset: null pointer"); return -1; }
      if (dry) return 0;
      if (cudaMemsetAsync(op.p[0], 0, n, st) != cudaSuccess) { set_error("memset: %s", cudaGetErrorString(cudaGetLastError())); return -3; }
      return 0;
    }
    case UCDIR_OP_TC_CONV: return launch_tc_conv(op, st, dry);
    case UCDIR_OP_TC_ATTN: return launch_tc_attn(op, st, dry);
    case UCDIR_OP_GN_APPLY_BF16: return launch_gn_apply_bf16(op, st, dry);
    case UCDIR_OP_CAST: return launch_cast(op, st, dry);
    case UCDIR_OP_CROP_TILES: return launch_crop_tiles(op, st, dry);
    case UCDIR_OP_GN_STATS_F32: return launch_gn_stats_f32(op, st, dry);
    case UCDIR_OP_GN_APPLY_F32: return launch_gn_apply_f32(op, st, dry);
    case UCDIR_OP_LAYOUT: return launch_layout(op, st, dry);
    case UCDIR_OP_TO_IMAGE_U8: return launch_to_image_u8(op, st, dry);
    default: set_error("unknown op kind %d", op.kind); return -1;
  }
}
}  // namespace ucdir

extern "C" {

int ucdir_run_ops(const ucdir_op_t* ops, int n_ops, void* stream) {
  if (!ops || n_ops < 0) { ucdir::set_error("run_ops: bad arguments"); return -1; }
  cudaStream_t st = (cudaStream_t)stream;
  if (ucdir::g_prof) ucdir::prof_mark(st, -1, 0);
  for (int k = 0; k < n_ops; ++k) {
    ucdir::NvtxOp range(ops[k], k);
    int rc = ucdir::dispatch(ops[k], st, false);
    if (ucdir::g_prof && !rc) ucdir::prof_mark(st, k, ops[k].kind);
    if (rc) { char tmp[400]; snprintf(tmp, sizeof(tmp), "%s", ucdir::g_err); ucdir::set_error("op %d (kind %d): %s", k, ops[k].kind, tmp); return rc; }
  }
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { ucdir::set_error("CUDA launch error: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return -3; }
  return 0;
}

int ucdir_check_ops(const ucdir_op_t* ops, int n_ops) {
  if (!ops || n_ops < 0) { ucdir::set_error("check_ops: bad arguments"); return -1; }
  for (int k = 0; k < n_ops; ++k) {
    int rc = ucdir::dispatch(ops[k], nullptr, true);
    if (rc) { char tmp[400]; snprintf(tmp, sizeof(tmp), "%s", ucdir::g_err); ucdir::set_error("op %d (kind %d): %s", k, ops[k].kind, tmp); return rc; }
  }
  return 0;
}

struct UcdirGraph { cudaGraph_t graph; cudaGraphExec_t exec; long long kernels; };

int ucdir_graph_capture(const ucdir_op_t* ops, int n_ops, void** graph_out) {
  if (!ops || n_ops <= 0 || !graph_out) { ucdir::set_error("graph_capture: bad arguments"); return -1; }
  if (ucdir::g_prof) { ucdir::set_error("graph_capture: not while profiling"); return -1; }
  cudaStream_t cs;
  if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) { ucdir::set_error("graph_capture: stream: %s", cudaGetErrorString(cudaGetLastError())); return -3; }
  const long long before = ucdir::g_launches;
  if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    ucdir::set_error("graph_capture: begin: %s", cudaGetErrorString(cudaGetLastError())); cudaStreamDestroy(cs); return -3; }
  int rc = 0;
  // UCDIR_OP_FLAG_BRANCH / JOIN: a side stream forked from (and joined back into) the capturing stream turns the flagged ops into
  // parallel branches of the graph (UCDIR_GRAPH_BRANCH=0: capture the list strictly in order)
  static const bool branches = []() { const char* e = getenv("UCDIR_GRAPH_BRANCH"); return !(e && e[0] == '0'); }();
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool pending = false;
  auto join = [&]() { if (pending) { cudaEventRecord(ev_join, side); cudaStreamWaitEvent(cs, ev_join, 0); pending = false; } };
  for (int k = 0; k < n_ops && !rc; ++k) {
    ucdir::NvtxOp range(ops[k], k);
    if (ops[k].flags & UCDIR_OP_FLAG_JOIN) join();
    cudaStream_t target = cs;
    if (branches && (ops[k].flags & UCDIR_OP_FLAG_BRANCH)) {
      if (!side && (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) != cudaSuccess)) {
        ucdir::set_error("graph_capture: side stream: %s", cudaGetErrorString(cudaGetLastError())); rc = -3; break; }
      join();                                         // one branch at a time
      cudaEventRecord(ev_fork, cs); cudaStreamWaitEvent(side, ev_fork, 0);
      target = side; pending = true;
    }
    rc = ucdir::dispatch(ops[k], target, false);
    if (rc) { char tmp[400]; snprintf(tmp, sizeof(tmp), "%s", ucdir::g_err); ucdir::set_error("graph_capture: op %d (kind %d): %s", k, ops[k].kind, tmp); }
  }
  join();
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(cs, &g);
  const long long kernels = ucdir::g_launches - before;
  ucdir::g_launches = before;                       // nothing ran yet; launches are counted per replay
  cudaStreamDestroy(cs);
  if (side) { cudaStreamDestroy(side); cudaEventDestroy(ev_fork); cudaEventDestroy(ev_join); }
  if (rc) { if (g) cudaGraphDestroy(g); return rc; }
  if (e != cudaSuccess || !g) { ucdir::set_error("graph_capture: end: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return -3; }
  cudaGraphExec_t ex = nullptr;
  e = cudaGraphInstantiate(&ex, g, 0);
  if (e != cudaSuccess) { ucdir::set_error("graph_capture: instantiate: %s", cudaGetErrorString(e)); cudaGraphDestroy(g); (void)cudaGetLastError(); return -3; }
  UcdirGraph* h = new UcdirGraph{g, ex, kernels};
  *graph_out = h;
  return 0;
}

int ucdir_graph_launch(void* graph, void* stream) {
  UcdirGraph* h = (UcdirGraph*)graph;
  if (!h) { ucdir::set_error("graph_launch: null graph"); return -1; }
  if (ucdir::g_nvtx) nvtxRangePushA("ucdir step graph");
  cudaError_t e = cudaGraphLaunch(h->exec, (cudaStream_t)stream);
  if (ucdir::g_nvtx) nvtxRangePop();
  if (e != cudaSuccess) { ucdir::set_error("graph_launch: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return -3; }
  ucdir::g_launches += h->kernels;
  return 0;
}

int ucdir_graph_destroy(void* graph) {
  UcdirGraph* h = (UcdirGraph*)graph;
  if (!h) return 0;
  cudaGraphExecDestroy(h->exec); cudaGraphDestroy(h->graph);
  delete h;
  return 0;
}

int ucdir_profile_begin(void) { ucdir::g_prof = true; ucdir::g_ev_used = 0; return 0; }

int ucdir_profile_end(float* ms, int* op_index, int* kind, int cap) {
  ucdir::g_prof = false;
  if (ucdir::g_ev_used == 0) return 0;
  if (cudaEventSynchronize(ucdir::g_ev[ucdir::g_ev_used - 1]) != cudaSuccess) { ucdir::set_error("profile_end: %s", cudaGetErrorString(cudaGetLastError())); return -3; }
  int n = 0;
  for (size_t k = 1; k < ucdir::g_ev_used && n < cap; ++k) {
    if (ucdir::g_ev_op[k] < 0) continue;                       // start marker of a later call
    float t = 0.f;
    cudaEventElapsedTime(&t, ucdir::g_ev[k - 1], ucdir::g_ev[k]);
    if (ms) ms[n] = t;
    if (op_index) op_index[n] = ucdir::g_ev_op[k];
    if (kind) kind[n] = ucdir::g_ev_kind[k];
    ++n;
  }
  ucdir::g_ev_used = 0;
  return n;
}

int ucdir_abi_version(void) { return UCDIR_ABI_VERSION; }
int ucdir_op_sizeof(void) { return (int)sizeof(ucdir_op_t); }
const char* ucdir_last_error(void) { return ucdir::g_err; }
long long ucdir_launch_count(void) { return ucdir::g_launches; }
int ucdir_tc_schedule(const ucdir_op_t* op) {
  if (!op || op->kind != UCDIR_OP_TC_CONV) return -1;
  if (ucdir::tc_final_halo_applies(*op)) return 3;
  if (ucdir::tc_mix_halo_applies(*op)) return 1;
  if (ucdir::tc_dense_halo_applies(*op)) return 2;
  return 0;
}

int ucdir_device_ok(void) {
  int dev = 0; cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    ucdir::set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError())); return -3; }
  if (prop.major != 10) { ucdir::set_error("device is sm_%d%d, this library is built for sm_100a only", prop.major, prop.minor); return -2; }
  return 0;
}

}  // extern "C"
