// Error channel + version of the C-ABI library.
#include <stdarg.h>
#include <stdlib.h>
#include "common.cuh"

static thread_local char g_err[1024] = "";

int tatt_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

extern "C" {
const char* tatt_last_error(void) { return g_err; }
int tatt_version(void) { return 100; }
// Node census of a captured CUDA graph (cudaGraph_t passed as void*): counts[0] kernel nodes, [1] memcpy, [2] memset,
// [3] everything else.  Host-only; used by bench.py to report the real number of kernel launches per replayed step.
int tatt_graph_node_counts(void* graph, int* counts) {
  size_t n = 0;
  TATT_CUDA(cudaGraphGetNodes((cudaGraph_t)graph, nullptr, &n));
  counts[0] = counts[1] = counts[2] = counts[3] = 0;
  if (n == 0) return 0;
  cudaGraphNode_t* nodes = (cudaGraphNode_t*)malloc(n * sizeof(cudaGraphNode_t));
  if (!nodes) return tatt_set_error("tatt_graph_node_counts: out of host memory");
  cudaError_t e = cudaGraphGetNodes((cudaGraph_t)graph, nodes, &n);
  for (size_t i = 0; e == cudaSuccess && i < n; ++i) {
    cudaGraphNodeType t;
    e = cudaGraphNodeGetType(nodes[i], &t);
    if (e != cudaSuccess) break;
    if (t == cudaGraphNodeTypeKernel) counts[0]++;
    else if (t == cudaGraphNodeTypeMemcpy) counts[1]++;
    else if (t == cudaGraphNodeTypeMemset) counts[2]++;
    else counts[3]++;
  }
  free(nodes);
  if (e != cudaSuccess) return tatt_set_error("tatt_graph_node_counts: %s", cudaGetErrorString(e));
  return 0;
}
int tatt_arch(void) {
#if defined(TATT_SM100A)
  return 1;
#else
  return 0;
#endif
}
}
