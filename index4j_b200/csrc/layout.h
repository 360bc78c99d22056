// Device-resident layout of an FmIndex (shared by the host re-layout code and the kernels).
// Everything the kernels fetch from HBM is a 32-byte, 32-byte-aligned record = one DRAM sector,
// read with a single 256-bit load (LDG.E.256).  DESIGN.md §3 explains each record.
#pragma once
#include <cstdint>

namespace fmgpu {

constexpr uint32_t SB_LOG = 20;                   // WaveletFixedBlockBoosting.SUPER_BLOCK_SIZE = 2^20
constexpr uint32_t SB_MASK = (1u << SB_LOG) - 1;
constexpr uint32_t SECTOR_BITS = 96;              // positions covered by one level record (3 words per bit plane)
constexpr uint32_t RRR_BLOCK = 15;                // RrrVector.BLOCK_SIZE
constexpr uint32_t SGROUP_BLOCKS = 32;            // RRR blocks per sampled-row group (480 bits)

struct alignas(32) Rec32 {
    uint32_t w[8];
};

// --- (block, symbol) cell, 8 bytes: WaveletFixedBlockBoosting.rank(pos, sym) for every position of the block, flattened
// (wavelet/WaveletFixedBlockBoosting.java:1022-1285).
//   value : hyper+super+block boundary rank (RUN / OCC_*) or the complete answer (CONST)
//   info  : kind << 29 | first occurrence record of the pair in `occ`
//   kinds : CONST   the symbol does not occur in the block (or the superblock): the answer is `value`, whatever the position
//                   (incl. the reference's scan for the next block that holds the symbol, :1048-1110, quirk Q3 restated)
//           RUN     single-symbol block: value + position in block (:1141-1146)
//           THROW   the reference indexes out of its arrays
//           OCC_*   the symbol occurs in the block: the reference walks the block's Huffman-shaped wavelet tree along the
//                   symbol's code (:1185-1279), one RRR rank per level, which counts the occurrences of the symbol among the
//                   first r positions of the block.  That count is stored directly, per (block, symbol), so a rank is the cell
//                   plus EXACTLY ONE record, whatever the code length:
//           every record: w0 = occurrences before the record (low 24 bits) | the symbol's code length << 24 (what the
//                     reference's walk costs: work counters only)
//           OCC_BITS  a bit vector over the block's positions: record q covers positions [224 q, 224 q + 224), w1..w7 = 224 bits
//           OCC_LIST / OCC_R4K / OCC_R1K  position lists over fixed position RANGES: record q holds the occurrences inside
//                     [q R, (q + 1) R) — at most 14 per range — as u16 positions (ascending, padded 0xffff) in w1..w7.
//                     OCC_LIST: R = 65536 = the whole block, i.e. ONE record for a symbol with <= 14 occurrences; R = 4096 /
//                     1024 for symbols that are neither rare nor frequent in a large block (typical of large alphabets):
//                     B / R records instead of B / 224 (a 456-symbol multi-script text: 3.2 x fewer occurrence records;
//                     1.3 % fewer for log text).  A pair with more than 14 occurrences in some 1024-range gets a bit vector.
// The table is dense, cells[block][symbol]: 8 bytes per pair keep it L2-resident for log alphabets (30 MB per 2^30 chars at 70
// symbols; it was 32 bytes per pair — 120 MB, which the 126 MB L2 did not hold beside the records: ncu showed the cell loads
// stalling as long as the DRAM-bound record loads — with in-cell position lists / list splitters that < 1 % of the queries used).
// (Rounds 1-2 kept the wavelet levels on this path too: one level record per two tree levels, up to three dependent records per
// rank, and a warp of 64 rank tracks in lockstep ran 2.9 record trips per step for 1.15 needed per track.  The level records
// remain what inverseSelect walks — the LF kernels — where the symbol is not known in advance.  Measured on the configs[1]
// batch, 65 % of the rank queries end in a bit vector, 32 % in RUN cells, 2.5 % in CONST cells, < 1 % in position lists.)
struct alignas(8) Cell8 {
    uint32_t value, info;
};
enum CellKind : uint32_t { CELL_NORMAL = 0, CELL_CONST = 1, CELL_RUN = 2, CELL_THROW = 3, CELL_OCC_R1K = 4, CELL_OCC_LIST = 5, CELL_OCC_BITS = 6, CELL_OCC_R4K = 7 };
constexpr uint32_t CELL_OCC_FIRST = CELL_OCC_R1K;  // kinds >= this one need their one occurrence record
constexpr uint32_t OCC_RANGE_MAX = 14;      // positions of a list record
// log2 of the position range a list record covers, by kind (OCC_R1K = 4: 10, OCC_LIST = 5: 16, OCC_R4K = 7: 12)
constexpr unsigned long long OCC_RANGE_SHIFTS = (10ull << (8 * 4)) | (16ull << (8 * 5)) | (12ull << (8 * 7));
constexpr uint32_t CELL_KIND_SHIFT = 29;
constexpr uint32_t CELL_PTR_MASK = (1u << CELL_KIND_SHIFT) - 1u;
constexpr uint32_t OCC_BITS_PER_REC = 224;   // positions per bit-vector record

// --- level record (32 bytes): TWO tree levels of 96 positions.  Only the wavelet-tree nodes at EVEN depth own records;
// record q of a node covers its positions [96q, 96q + 96):
//   w0 = c00 | c01 << 16, w1 = c10 | c11 << 16 : elements before the record whose bits at this level and the next are
//            (t, u) = (0,0), (0,1), (1,0), (1,1)  (node-relative, 16 bits each)
//   w2..w4 = plane 0: this node's 96 bits
//   w5..w7 = plane 1: for the SAME 96 positions the bit each element has one level further down, i.e. in the child it
//            descends to (0 where that child is a leaf)
// The rank of any 2-bit code prefix (t, u) is ONE fetch, one 16-bit field and one masked-popcount pass over
// (plane0 ^ ~T) & (plane1 ^ ~U): a rank / inverseSelect takes one record per TWO tree levels.  A code that ends at the
// record's first level is the case u = 0 (plane 1 is 0 for elements that go to a leaf child), so there is no special case.
// (v1 kept one 224-bit level per 32-byte sector: one DRAM access per level, and the popcount passes over 7 words made the
// two-level variant of that format ALU-bound — profiles/experiments.)
// Every node starts on a fresh record, so the counters are node-relative (nodes hold <= 65536 elements).

// --- block descriptor (LF / inverseSelect entry, :1305-1537):
//   w0 first level record of the root, w1 info (bit0 = run block, bits 8-23 = run symbol c' as inverseSelect decodes it),
//   run blocks: w2 / w3 = value / kind of the (block, c') cell, i.e. rank(j, c') for j inside the block;
//   w4 = node record of the root; run blocks: w5 = the boundary ranks inverseSelect adds the in-block position to (:1338-1352).
// --- node record of an even-depth node (32 B): entries [t][u] = {c, a} at w[4t + 2u], t = bit at this depth, u = bit at
// the next.  Child t is a leaf: entry [t][0] = {LEAF1_FLAG | symbol, boundary rank of the symbol}.  Else grandchild [t][u]:
// leaf => {LEAF_FLAG | symbol, boundary rank}; internal => {its node record, its first level record}.
constexpr uint32_t LEAF_FLAG = 0x80000000u;
constexpr uint32_t LEAF1_FLAG = 0x40000000u;

// --- sampled-row group (RrrVector restated for one 32-block group, bitsequence/RrrVector.java:314-396):
//   w0 ones before the group, w1 bit position of the group's first offset in the offset stream,
//   w2 = cumulative offset bits at block 8,16,24 (3 x 10 bits), w3 = cumulative ones at block
//   8,16,24 (3 x 10 bits), w4..w7 = 32 class nibbles (block k = nibble k).

struct SbDesc {
    uint32_t first_block;  // global block number of the superblock's block 0
    uint32_t block_log;    // blockSizeLog
};

struct U32x2 {
    uint32_t x, y;
};

struct PatDesc {  // one per pattern, written by the pre-pass
    uint64_t off;     // offset of the pattern's first char in the concatenated code array
    uint32_t len;
    uint32_t last;    // code of the last char (the first one backward search consumes), or PAT_KMER | q-gram table index
};
constexpr uint32_t PAT_KMER = 0x80000000u;

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
    uint32_t reserved0;
    // q-gram start table of the backward search (0 = none): for every q-gram of alphabet codes the SA range after its q chars,
    // i.e. the state of FmIndex.count after q - 1 steps, computed at load by the search kernel itself.  Entry of the q-gram
    // whose LAST char has code a, the one before b, ... : index ((a * stride + b) * stride + ...); {0xffffffff, .} = not usable
    // (a step of that q-gram throws in the reference), the pattern then starts from its last char as usual.
    uint32_t kmer_q;
    uint32_t kmer_stride;
    const U32x2* kmer;
    const uint32_t* C;
    const uint16_t* char2code;  // [65536], 0 = not in alphabet (monotonicMap.getOrDefault(c, 0))
    const uint16_t* code2char;  // monotonicLookUp
    const SbDesc* sb;
    const Cell8* cells;    // [n_blocks_total][sigma]
    const Rec32* sectors;  // level records of the even-depth wavelet nodes (inverseSelect)
    const Rec32* occ;      // occurrence records of the (block, symbol) pairs (position lists / bit vectors)
    const Rec32* blocks;   // block descriptors
    const Rec32* nodes;    // node records of the even-depth nodes
    const Rec32* sgroups;
    const uint32_t* soffsets;
    const uint16_t* rrr_inv;    // [32768] (class, offset) -> 15-bit block, RrrVector.java:8705-16899 regenerated
    const uint16_t* rrr_cbase;  // [16]    RrrVector.java:8692-8698
    const Rec32* sa;       // SA samples, 8 per Rec32
    const Rec32* isa;      // ISA samples, 8 per Rec32
    // Device-side DENSER sampling of the SA rows for locate (0 = none; fmgpu.cu build_dense_samples): the serialized index
    // samples the rows whose suffix starts at a multiple of sampleRate (fm/FmIndex.java:343-357); at load the LF kernels walk
    // the text once and also mark the rows of every multiple of dense_rate (a divisor of sampleRate), so a hit's walk
    // (fm/FmIndex.java:531-537) ends after (dense_rate - 1) / 2 steps on average instead of (sampleRate - 1) / 2 — the same
    // positions, HBM traded for LF steps.
    //   dmarks record q: w0 = marked rows before row 224 q, w1..w7 = marks of rows [224 q, 224 q + 224)
    //   dsa[k] = suffix start of the k-th marked row
    uint32_t dense_rate;
    uint32_t n_dense;
    const Rec32* dmarks;
    const uint32_t* dsa;
};
constexpr uint32_t DENSE_ROWS_PER_REC = 224;

}  // namespace fmgpu
