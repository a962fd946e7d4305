#pragma once
#include <cuda_runtime.h>
#include <mutex>

// Per-device, once-only kernel setup (cudaFuncSetAttribute is per device; a process may drive several GPUs through
// one handle each).  One table per kernel; get() runs `init` the first time the CURRENT device asks.
struct PerDeviceInt {
  static constexpr int kMaxDevices = 64;
  int value[kMaxDevices] = {};
  bool set[kMaxDevices] = {};
  std::mutex mu;
  template <typename F>
  int get(F init) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) return init();
    std::lock_guard<std::mutex> lock(mu);
    if (!set[dev]) { value[dev] = init(); set[dev] = true; }
    return value[dev];
  }
};

