// Layout plan of one weight-norm MLP packed for the tcgen05 tile core.
//
// Packed blob (device memory, owned by the caller):
//   for every layer, for every 128-row output tile m, for every 32-wide K chunk kc:
//       [ hi tile : 128 x 32 fp16 ][ lo tile : 128 x 32 fp16 ]          = 16 KiB = one pipeline stage
//   tile element (r, k) at (r/8)*512 + (k/8)*128 + (r%8)*16 + (k%8)*2   (UMMA K-major, SWIZZLE_NONE)
//   W_eff * kWeightScale = hi + lo   (fp16 two-term split, ~22 significant bits)
//   followed by the fp32 bias area (one float per padded output row).
// The reference computes the same layers with fp32 cuBLAS SGEMMs
// (implicit_differentiable_renderer.py:89 and :160).
#pragma once
#include <cstdint>

namespace mvsdf {

constexpr int kTileM = 128;        // output features per UMMA (TMEM lanes)
constexpr int kTileN = 64;         // points (columns) per tile
constexpr int kChunkK = 32;        // K elements per pipeline stage
#ifndef MVSDF_STAGES
#define MVSDF_STAGES 4
#endif
constexpr int kStages = MVSDF_STAGES;
constexpr int kTileBytes = kTileM * kChunkK * 2;   // 8 KiB (hi or lo)
constexpr int kStageBytes = 2 * kTileBytes;        // 16 KiB
// Activation operand (B of the UMMA), MN-major / SWIZZLE_NONE: element (column n, feature k) lives at
//   (k/8)*kBCoreStride + (n/8)*128 + (k%8)*16 + (n%8)*2
// i.e. per block of 8 features: 16 column blocks of 128 B -- blocks 0-7 hold the fp16 hi parts of the 64 columns,
// blocks 8-15 the lo parts (one N=128 descriptor covers [X_hi ; X_lo]).  A thread of the epilogue owns one feature and
// consecutive columns, so its outputs are contiguous 16-byte vectors.
constexpr int kBLoOffset = 8 * 128;                // byte offset of the lo half inside a feature block
constexpr int kBCoreStride = 16 * 128;             // 2048: byte stride between blocks of 8 features (the descriptor's LBO)
constexpr float kWeightScale = 64.0f;              // power of two: keeps the lo parts out of fp16 subnormals
constexpr float kActScale = 64.0f;
constexpr int kMaxLayers = 12;
constexpr int kTmemCols = 512;                     // 4 output tiles x (64 + 64) accumulator columns

// status words at the end of the packed blob (zeroed by every mvsdf_pack_weights)
constexpr int kStatusWords = 4;
constexpr int kStatusPackRange = 0;   // packed weight elements with |kWeightScale * W| beyond the fp16 range (or non-finite)
constexpr int kStatusNonFinite = 1;   // non-finite head outputs written by the MLP kernels (an activation left the fp16 range)

enum Act : int { ACT_NONE = 0, ACT_SOFTPLUS100 = 1, ACT_RELU = 2 };
enum NetKind : int { NET_SDF = 0, NET_RENDER = 1 };
enum HeadKind : int { HEAD_SDF_ONLY = 0, HEAD_FULL = 1 };

struct LayerPlan {
  int in_dim;        // logical input width of the source weight
  int out_dim;       // logical output rows of the source weight
  int k_chunks;      // padded K / 32
  int m_tiles;       // padded rows / 128
  int act;
  int b_from_pe;     // B operand = positional-encoding tile (first SDF layer)
  int bias_off;      // float index into the bias area
  int row_map;       // 0 identity, 1 head_sdf (row0 only), 2 head_full (features first, then sdf+indicator)
  int src_layer;     // index of the source lin{l} this plan entry is packed from
  long long w_off;   // byte offset of tile (m=0, kc=0)
  float col_scale;   // extra scale folded into the weights (1/sqrt(2) at the skip layer)
};

struct NetPlan {
  int kind;
  int width;
  int n_hidden;          // layers with an activation
  int n_layers;          // entries in L used by the packer (hidden + heads)
  int skip_layer;        // layer whose input is cat([h, PE])/sqrt(2); -1 if none
  int skip_rows_begin;   // first feature row that holds PE at the skip layer input (out_dim of layer skip-1)
  int pe_dim;            // 39 for the SDF net
  int head_index[2];     // L index of HEAD_SDF_ONLY / HEAD_FULL (render: both = last layer)
  int k_cores_max;       // max over layers of padded K / 8   (activation buffer size)
  int feat_size;
  long long bias_area_off;
  long long scale_area_off;   // fp32 g/||v|| per source row, scratch of the packer
  long long status_off;       // int32[kStatusWords]: range / non-finite monitors (see kStatus*)
  int n_src_layers;
  int scale_off[kMaxLayers];  // float index per source layer
  long long total_bytes;
  LayerPlan L[kMaxLayers];
};

}  // namespace mvsdf
