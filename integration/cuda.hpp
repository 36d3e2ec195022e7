// integration/cuda.hpp — drop-in replacement for the reference's src/xpu/cuda.hpp (the empty GPU
// device stub, reference src/xpu/cuda.hpp:8-13): `cuda_t : xpu_t` backed by libphos_cuda.so.
// Compiled against the reference's own headers by `make -C oracle ref` (which copies this file over
// src/xpu/cuda.hpp of its temporary build copy); see INTEGRATION.md.
#pragma once

#include "xpu.hpp"

#include <atomic>
#include <cstddef>
#include <cstdint>
#include <exception>
#include <thread>

struct parsed_options_t;
struct phos_ctx;

/* Implements logic for NVIDIA GPUs (B200, sm_100a) via libphos_cuda.so */
struct cuda_t : public xpu_t {
  phos_ctx* ctx = nullptr;
  std::thread worker;
  std::exception_ptr error;  // what the worker threw; rethrown by join()
  uint32_t spp = 16;
  uint64_t seed = 0;
  uint32_t film_w = 0, film_h = 0;
  float* pinned = nullptr;  // page-locked landing zone of the film rows of a chunk
  size_t pinned_bytes = 0;
  uint32_t tiles_done = 0, claims = 0;  // of the last frame (tests / diagnostics)
  static std::atomic<int> instances;    // live cuda_t devices: the peers that share a frame's tile queue

  ~cuda_t();

  void preprocess(const scene_t& scene) override;
  void start(const scene_t& scene, frame_state_t& state) override;
  void join() override;

  static int device_count();
  static cuda_t* make(const parsed_options_t& options, int device = 0);
};
