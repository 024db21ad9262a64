// Layout contract between the weight packer and the fused render kernel.
//
// The NeRF_sigma MLP (reference models/nerf.py:137-154: 8 x Linear(256)+ReLU with
// the embedded xyz re-concatenated in front of the hidden state at layer 5, a
// 256->1 softplus sigma head, Linear 256->256, [final|dir] -> 128 ReLU, 128 -> 64
// sigmoid) is executed per 128-point tile as a fixed PROGRAM of tensor-core
// "units"; each unit is one (layer, 128-wide output half) accumulated over a list
// of weight CHUNKS.  A chunk is a [rows x 64] K-major, 128-byte-swizzled slab of
// 16-bit weights - exactly the shared-memory image a tcgen05.mma B descriptor
// reads - so the kernel streams the packed image with plain bulk copies.
//
// A-operand sources of a chunk:
//   kSrcEmb : the tile's embedding buffer in shared memory (SS MMA). 128 columns:
//             [0,e_xyz) xyz embedding, zero pad to 96, [96,96+e_dir) dir embedding, zero
//             pad to 128; two SW128 slabs of 64 columns.
//   kSrcAct : the previous layer's activations in TMEM (TS MMA), 2 values per column.
//
// Biases stay in fp32: the packer copies them into the side blob and the layer
// epilogue adds them to the drained accumulator before the activation.  (An earlier
// revision ran them through the tensor core as an extra 16-column k-step per unit;
// that cost one ring slot, one weight copy, one MMA and one commit per unit - the
// weight ring then held only 1.6 units and the issuer waited on it.  With pure
// weight chunks a standard unit is exactly four 16 KB slabs and the 8-slot ring holds
// two whole units.)
//
// build_program() is the single source of truth.  The host runs it per call (microseconds); the
// pack kernel receives the tables as a kernel parameter and stores them BEHIND the weight image
// and the fp32 blob, so every packed buffer carries its own program: the render kernel reads the
// tables from the buffer it was given - no __constant__ symbol, no upload, no host
// synchronisation, no state shared between devices, streams or embedding widths.
//
// Operand formats (crnerf_operand): 0 fp16, 1 bf16 - one 16-bit image; 2 "fp16x3" - two images,
// W_hi = fp16(W) and W_lo = fp16(W - W_hi), with activations split the same way in the kernel
// and every product issued as three MMAs (hi*hi + lo*hi + hi*lo): fp32-class accuracy at three
// times the tensor work, for weights whose cancellation the 11-bit operands cannot hold to 1e-4.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CRNERF_HD __host__ __device__
#else
#define CRNERF_HD
#endif

namespace crnerf {

constexpr int kWidth = 256;      // trunk width W
constexpr int kDepth = 8;        // trunk depth D
constexpr int kSkipLayer = 4;    // skips=[4]
constexpr int kDirWidth = 128;   // W/2
constexpr int kOutDim = 64;      // nerf_out_dim
constexpr int kEmbCols = 128;    // embedding buffer columns (2 slabs of 64)
constexpr int kDirCol0 = 96;     // first dir-embedding column in the buffer
constexpr int kMaxExyz = 96, kMaxEdir = 32;

enum LayerId : int {
  kL1 = 0,  // .. kL8 = 7
  kLFinal = 8,
  kLDir = 9,
  kLRgb = 10,
  kNumLayers = 11,
  kLSigma = 11  // not a tensor-core layer; index into the weight pointer array only
};
enum ASrc : int { kSrcEmb = 0, kSrcAct = 1 };
enum ChunkKind : int { kKindWeights = 0 };  // (kind 1 was the bias chunk of earlier revisions)

struct Chunk {
  int32_t offset;   // byte offset in the packed image
  int32_t bytes;    // rows * 128
  int16_t layer;    // LayerId
  int16_t rows;     // output neurons in this chunk (128 or 64)
  int16_t row0;     // first output neuron
  int16_t wcol0;    // column of the reference weight matrix mapped to K index 0
  int16_t wcols;    // valid K columns (<= 64); the rest of the slab is zero
  int16_t a_src;    // ASrc
  int16_t a_k0;     // first k-step (16 columns) in the A source
  int16_t nk;       // k-steps to issue = ceil(wcols / 16)
  int16_t kind;     // ChunkKind
  int16_t pad;
};

struct Unit {
  int16_t layer;        // LayerId
  int16_t half;         // output half (0/1) for 256-wide layers, 0 otherwise
  int16_t n;            // MMA N (128 or 64)
  int16_t chunk0;       // first chunk index
  int16_t nchunks;
  int16_t first_of_layer;
  int16_t last_of_layer;
  int16_t pad;
};

constexpr int kMaxChunks = 104;
constexpr int kMaxUnits = 20;

struct Program {
  Chunk chunks[kMaxChunks];
  Unit units[kMaxUnits];
  int32_t n_chunks;
  int32_t n_units;
  int32_t image_bytes;
  int32_t e_xyz, e_dir;
};

// tables stored behind the blob in every packed buffer (see the header comment)
struct Tables {
  Chunk chunks[kMaxChunks];
  Unit units[kMaxUnits];
  int32_t n_chunks;
  int32_t n_units;
  int32_t image_bytes;
  int32_t n_images;   // 1, or 2 for the split (fp16x3) format
};

// fp32 side blob: sigma head weights (256) + sigma bias, then the biases of the 11
// tensor-core layers in LayerId order (9 x 256, 128, 64).  The kernel stages the whole blob
// (11 KB) in shared memory: bias reads through L1 miss too often to sit in the epilogue.
constexpr int kSigmaWOff = 0;
constexpr int kSigmaBOff = 256;
constexpr int kBiasOff = 264;
constexpr int kBlobFloats = kBiasOff + 9 * 256 + 128 + 64;  // 2760
CRNERF_HD inline int bias_offset(int layer) {  // float offset of a tensor-core layer's bias in the blob
  return kBiasOff + (layer <= 8 ? 256 * layer : (layer == 9 ? 9 * 256 : 9 * 256 + 128));
}

// Index of each layer's tensors in the caller-provided pointer arrays
// (crnerf_mlp_weights_t): xyz_encoding_1..8, xyz_encoding_final, dir_encoding,
// static_rgb, static_sigma.
CRNERF_HD inline int layer_in_features(int layer, int e_xyz, int e_dir) {
  if (layer == 0) return e_xyz;
  if (layer == kSkipLayer) return e_xyz + kWidth;
  if (layer < 8 || layer == kLFinal || layer == kLSigma) return kWidth;
  if (layer == kLDir) return kWidth + e_dir;
  return kDirWidth;  // rgb
}
CRNERF_HD inline int layer_out_features(int layer) {
  if (layer < 8 || layer == kLFinal) return kWidth;
  if (layer == kLDir) return kDirWidth;
  if (layer == kLRgb) return kOutDim;
  return 1;
}

constexpr int kNumOperands = 3;
CRNERF_HD inline int operand_images(int operand) { return operand == 2 ? 2 : 1; }
CRNERF_HD inline int operand_fmt(int operand) { return operand == 1 ? 1 : 0; }   // 16-bit element format

inline void build_program(int e_xyz, int e_dir, Program* p) {
  int nc = 0, nu = 0, off = 0;
  auto add_chunk = [&](int layer, int row0, int rows, int wcol0, int wcols, int a_src, int a_k0) {
    if (wcols <= 0) return;
    Chunk& c = p->chunks[nc++];
    c.offset = off;
    c.bytes = rows * 128;
    c.layer = (int16_t)layer;
    c.rows = (int16_t)rows;
    c.row0 = (int16_t)row0;
    c.wcol0 = (int16_t)wcol0;
    c.wcols = (int16_t)wcols;
    c.a_src = (int16_t)a_src;
    c.a_k0 = (int16_t)a_k0;
    c.nk = (int16_t)((wcols + 15) / 16);
    c.kind = kKindWeights;
    c.pad = 0;
    off += c.bytes;
  };
  auto emb_chunks = [&](int layer, int row0, int rows, int wcol0, int cols, int bufcol0) {
    // cols embedding columns starting at buffer column bufcol0, split at slab (64) boundaries
    int done = 0;
    while (done < cols) {
      int bc = bufcol0 + done;
      int take = 64 - (bc % 64);
      if (take > cols - done) take = cols - done;
      add_chunk(layer, row0, rows, wcol0 + done, take, kSrcEmb, bc / 16);
      done += take;
    }
  };
  auto act_chunks = [&](int layer, int row0, int rows, int wcol0, int cols) {
    for (int j = 0; j < cols; j += 64) add_chunk(layer, row0, rows, wcol0 + j, 64, kSrcAct, j / 16);
  };
  auto begin_unit = [&](int layer, int half, int n, int first, int last) {
    Unit& u = p->units[nu];
    u.layer = (int16_t)layer;
    u.half = (int16_t)half;
    u.n = (int16_t)n;
    u.chunk0 = (int16_t)nc;
    u.first_of_layer = (int16_t)first;
    u.last_of_layer = (int16_t)last;
    u.pad = 0;
  };
  auto end_unit = [&]() {
    Unit& u = p->units[nu];
    u.nchunks = (int16_t)(nc - u.chunk0);
    nu++;
  };
  for (int l = 0; l < kDepth; ++l) {
    for (int h = 0; h < 2; ++h) {
      begin_unit(l, h, 128, h == 0, h == 1);
      if (l == 0) {
        emb_chunks(l, h * 128, 128, 0, e_xyz, 0);
      } else if (l == kSkipLayer) {
        emb_chunks(l, h * 128, 128, 0, e_xyz, 0);
        act_chunks(l, h * 128, 128, e_xyz, kWidth);
      } else {
        act_chunks(l, h * 128, 128, 0, kWidth);
      }
      end_unit();
    }
  }
  for (int h = 0; h < 2; ++h) {
    begin_unit(kLFinal, h, 128, h == 0, h == 1);
    act_chunks(kLFinal, h * 128, 128, 0, kWidth);
    end_unit();
  }
  begin_unit(kLDir, 0, 128, 1, 1);
  act_chunks(kLDir, 0, 128, 0, kWidth);
  emb_chunks(kLDir, 0, 128, kWidth, e_dir, kDirCol0);
  end_unit();
  begin_unit(kLRgb, 0, 64, 1, 1);
  act_chunks(kLRgb, 0, 64, 0, kDirWidth);
  end_unit();
  p->n_chunks = nc;
  p->n_units = nu;
  p->image_bytes = off;
  p->e_xyz = e_xyz;
  p->e_dir = e_dir;
}

}  // namespace crnerf
