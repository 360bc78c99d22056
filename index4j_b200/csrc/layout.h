// Device-resident layout of an FmIndex (shared by the host re-layout code and the kernels).
// Everything the kernels fetch from HBM is a 32-byte, 32-byte-aligned record = one DRAM sector,
// read with a single 256-bit load (LDG.E.256).  DESIGN.md §3 explains each record.
#pragma once
#include <cstdint>

namespace fmgpu {

constexpr uint32_t SB_LOG = 20;                   // WaveletFixedBlockBoosting.SUPER_BLOCK_SIZE = 2^20
constexpr uint32_t SB_MASK = (1u << SB_LOG) - 1;
constexpr uint32_t SECTOR_BITS = 224;             // payload bits of a level sector (7 words)
constexpr uint32_t RRR_BLOCK = 15;                // RrrVector.BLOCK_SIZE
constexpr uint32_t SGROUP_BLOCKS = 32;            // RRR blocks per sampled-row group (480 bits)

struct alignas(32) Rec32 {
    uint32_t w[8];
};

// --- (block, symbol) cell: everything WaveletFixedBlockBoosting.rank(pos, sym) reads besides the
// level bits, flattened (wavelet/WaveletFixedBlockBoosting.java:1022-1184).
//   w0 value : hyper+super+block boundary rank (NORMAL/RUN) or the complete answer (CONST)
//   w1 code  : canonical Huffman code of sym in this block, MSB = root decision
//   w2       : codeLen (bits 0-7) | kind (bits 8-15)
//   w3..w7   : sector index of the node visited at depth 0..4; when codeLen > 5, w3..w6 hold depth
//              0..3 and w7 is an index into the overflow array (chunks of 8 sector indices, depth 4..)
enum CellKind : uint32_t { CELL_NORMAL = 0, CELL_CONST = 1, CELL_RUN = 2, CELL_THROW = 3 };
constexpr uint32_t CELL_INLINE_LEVELS = 5;

// --- level sector: w0 = ones in this node's bitvector before the sector, w1..w7 = 224 bits.
// Every wavelet-tree node starts on a fresh sector, so w0 is node-relative and a rank inside a
// node touches exactly one sector.

// --- block descriptor (LF / inverseSelect entry, :1305-1537):
//   w0 root sector, w1 info (bit0 = run block, bits 8-23 = run symbol c' as inverseSelect decodes it),
//   run blocks: w2 / w3 = value / kind of the (block, c') cell, i.e. rank(j, c') for j inside the block;
//   w4..w7 = node record of the root.
// --- node record (16 B, two per sector): c0, c1, a0, a1.  Child b: c_b bit31 set => leaf, low 16
// bits = symbol, a_b = boundary rank of that symbol; else c_b = record index of the child node
// and a_b = its first sector.
constexpr uint32_t LEAF_FLAG = 0x80000000u;

// --- sampled-row group (RrrVector restated for one 32-block group, bitsequence/RrrVector.java:314-396):
//   w0 ones before the group, w1 bit position of the group's first offset in the offset stream,
//   w2 = cumulative offset bits at block 8,16,24 (3 x 10 bits), w3 = cumulative ones at block
//   8,16,24 (3 x 10 bits), w4..w7 = 32 class nibbles (block k = nibble k).

struct SbDesc {
    uint32_t first_block;  // global block number of the superblock's block 0
    uint32_t block_log;    // blockSizeLog
};

struct PatDesc {  // one per pattern, written by the pre-pass
    uint64_t off;     // offset of the pattern's first char in the concatenated code array
    uint32_t len;
    uint32_t last;    // code of the last char (the first one backward search consumes)
};

struct DevIndex {
    // FmIndex scalars
    uint32_t length;       // n + 1
    uint32_t sample_rate;
    uint32_t n_c;          // entries of C
    uint32_t n_lookup;
    uint32_t sigma;        // wavelet alphabet size
    uint32_t n_sb;
    uint32_t q4;           // 1 if length % 2^20 == 0 (rank(size, .) throws in the reference)
    uint32_t extract_enabled;
    uint32_t n_isa;        // entries of positions (ISA samples)
    uint32_t n_sa;
    uint32_t s_total_ones;
    uint32_t pad0;
    const uint32_t* C;
    const uint16_t* char2code;  // [65536], 0 = not in alphabet (monotonicMap.getOrDefault(c, 0))
    const uint16_t* code2char;  // monotonicLookUp
    const SbDesc* sb;
    const Rec32* cells;    // [n_blocks_total][sigma]
    const Rec32* sectors;
    const Rec32* ovf;      // chunks of 8 sector indices
    const Rec32* blocks;   // block descriptors
    const Rec32* nodes;    // node records, two per Rec32
    const Rec32* sgroups;
    const uint32_t* soffsets;
    const uint16_t* rrr_inv;    // [32768] (class, offset) -> 15-bit block, RrrVector.java:8705-16899 regenerated
    const uint16_t* rrr_cbase;  // [16]    RrrVector.java:8692-8698
    const Rec32* sa;       // SA samples, 8 per Rec32
    const Rec32* isa;      // ISA samples, 8 per Rec32
};

}  // namespace fmgpu
