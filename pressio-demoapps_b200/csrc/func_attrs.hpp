// func_attrs.hpp -- per-device kernel attribute set-up.  cudaFuncSetAttribute acts on the CURRENT device's context, so
// a "configured once per process" flag leaves a second device of the same process (same-process slab neighbours,
// pda_slab_peer_connect_local) launching with more than 48 KB of dynamic shared memory unconfigured.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <string>
#include <utility>

#include "common.hpp"

namespace pda {

template <class K>
inline void ensureFuncAttrs(K kern, int smemBytes, bool maxCarveout = false) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;   // (kernel, device) -> configured dynamic smem bytes
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) throw Error(kCuda, std::string("cudaGetDevice failed: ") + cudaGetErrorString(e));
  const std::pair<const void*, int> key{reinterpret_cast<const void*>(kern), dev};
  std::lock_guard<std::mutex> lk(mu);
  auto it = done.find(key);
  if (it != done.end() && it->second >= smemBytes) return;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
  if (e == cudaSuccess && maxCarveout) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) throw Error(kCuda, std::string("cudaFuncSetAttribute failed: ") + cudaGetErrorString(e));
  done[key] = smemBytes;
}

}  // namespace pda
