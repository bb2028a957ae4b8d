// sart_internal.h — private to libsart.so: the handle, error reporting, launcher prototypes.
#pragma once
#include <cstdarg>
#include <cstddef>
#include <cstdint>

#include "../../include/sart.h"
#include "device_params.h"

typedef struct CUstream_st* cudaStream_t;

namespace sart {

// Records the message for sart_last_error() and returns `code`.
int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));

// Host derivation of the kernel parameter blocks (derive.cpp).
void derive_params(const sart_setup_t& s, const sart_tables_t* tb, Params* p);
void derive_shells(const sart_setup_t& s, ShellF64* shells /* [SART_MAX_SHELLS] */);

}  // namespace sart

struct sart_handle {
  int device = -1;
  cudaStream_t stream = nullptr;
  sart_setup_t setup;
  sart::Params params;
  sart::Tables tables;          // device pointers
  void* table_blob = nullptr;   // one allocation backing every table
  size_t table_bytes = 0;
  int precision = 0;            // 0 exact f64, 1 fast
  int sm_count = 148;
  size_t shell_offset = 0;      // byte offset of the ShellF64 array inside table_blob
  int n_masses = 1;
  int masses_default = 1;       // masses[0] still mirrors setup.consts.mAxion
  int image_masses = 0;         // mass planes the image buffers were allocated for
  double masses[SART_MAX_MASSES] = {0};
  double* d_masses = nullptr;   // [SART_MAX_MASSES]
  double* d_image = nullptr;    // [n_masses][256][256]
  double* d_image_w2 = nullptr;
  sart_counters_t* d_counters = nullptr;  // [n_masses]
  // staging for the host-pointer entry points
  void* d_stage = nullptr;
  size_t stage_bytes = 0;
  void* h_stage = nullptr;      // pinned
  size_t h_stage_bytes = 0;
};
