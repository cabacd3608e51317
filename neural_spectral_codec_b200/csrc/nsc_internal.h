// Internal declarations shared by the translation units of libnsc_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nsc_b200.h"

namespace nsc {

constexpr int kAz = NSC_N_AZIMUTH;      // 360 columns
constexpr int kFreqs = NSC_N_FREQS;     // 181
constexpr int kPitch = kAz + 1;         // smem row pitch of the min-image: column 360 catches az == 2*pi
constexpr int kMaskWords = 12;          // 360 validity bits per row
constexpr uint32_t kInfBits = 0x7f800000u;
constexpr int kColTerms = 7;            // odd polynomial for atan on [0,1] (azimuth)
constexpr int kRowTerms = 5;            // odd polynomial for atan on the elevation FOV
constexpr int kMaxSignals = 8;          // complex 360-point FFTs in flight per CTA (two rows each)

enum RowMode { kRowPoly = 0, kRowSearch = 1 };

// Kernel arguments derived on the host from nsc_params (+ the freq->bin table). Passed by
// value as a __grid_constant__ so every field is a constant-bank operand.
struct DeviceParams {
    int E;                 // projected rows
    int T;                 // target rows after pooling
    int n_bins;
    int interpolate;
    int row_mode;          // RowMode
    float s_lo, s_hi;      // keep a point iff s_lo <= (x*x + y*y) + z*z <= s_hi (host_tables.cu)
    float eps;
    float uniform;         // 1 / (T * n_bins) in float32
    float row_p[kRowTerms];   // (atan(u) - el_min) / row_width = row_off + u * P(u^2)
    float row_off;
    float u_lo, u_hi;      // clamp of u = z / rho that keeps the row value inside (0, E)
    float row_c[NSC_MAX_ELEVATION];  // kRowSearch: row_c[k] = tan(edge_k)*|tan(edge_k)|, k = 1..E-1
    uint8_t bin_start[NSC_MAX_BINS + 3];  // first frequency of bin b; bin_start[n_bins] = 181
};

// Fills DeviceParams; returns an nsc_status. Results are memoised per (params, lut).
int make_device_params(const nsc_params* p, const int32_t* h_lut, DeviceParams* out);
int validate_params(const nsc_params* p);

// thread-local CUDA error text for nsc_last_cuda_error().
int record_cuda(cudaError_t e);

// Launchers (nsc_encode.cu)
int launch_encode(const float* d_points, int stride, const long long* d_offsets, long long origin,
                  int n_scans, const DeviceParams& dp, float* d_out, float* d_img_out, int stage,
                  float* const* d_peer_out, int n_peers, long long peer_row0,
                  unsigned* d_workspace, size_t workspace_bytes, cudaStream_t stream);
int launch_encode_images(const float* d_images, int n_images, int rows, const DeviceParams& dp,
                         float* d_out, cudaStream_t stream);
int launch_interpolate(const float* d_in, int n_images, int rows, int nearest, float* d_out, cudaStream_t stream);
size_t workspace_bytes_for(int n_scans, int E);   // recommended: room to split the last wave's scans
size_t workspace_bytes_min();                      // enough for every launch (no splitting)

}  // namespace nsc
