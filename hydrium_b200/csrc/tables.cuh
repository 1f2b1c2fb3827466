// hydrium_b200/csrc/tables.cuh
//
// Constant tables of the fixed quantiser / scan order / context model
// (reference: encoder.c:32-95).  Numeric tables are data of the JPEG XL encoder configuration
// the reference implements (BSD-2-Clause, Leo Izen); layout and use here are our own.
#pragma once

#include "common.cuh"

namespace hydb {

// cosine_lut (encoder.c:32-40) as IEEE-754 bit patterns: the reference's literals are decimal
// doubles rounded to float by its compiler; SURVEY.md Appendix B.3 lists the resulting bits.
#define HYDB_COS_BITS                                                                                   \
    {0x3e318a87u, 0x3e1682f9u, 0x3dc92352u, 0x3d0d42a9u, 0xbd0d42a9u, 0xbdc92352u, 0xbe1682f9u, 0xbe318a87u}, \
    {0x3e273d5cu, 0x3d8a8bd2u, 0xbd8a8bd2u, 0xbe273d5cu, 0xbe273d5cu, 0xbd8a8bd2u, 0x3d8a8bd2u, 0x3e273d5cu}, \
    {0x3e1682f9u, 0xbd0d42a9u, 0xbe318a87u, 0xbdc92352u, 0x3dc92352u, 0x3e318a87u, 0x3d0d42a9u, 0xbe1682f9u}, \
    {0x3e000000u, 0xbe000000u, 0xbe000000u, 0x3e000000u, 0x3e000000u, 0xbe000000u, 0xbe000000u, 0x3e000000u}, \
    {0x3dc92352u, 0xbe318a87u, 0x3d0d42a9u, 0x3e1682f9u, 0xbe1682f9u, 0xbd0d42a9u, 0x3e318a87u, 0xbdc92352u}, \
    {0x3d8a8bd2u, 0xbe273d5cu, 0x3e273d5cu, 0xbd8a8bd2u, 0xbd8a8bd2u, 0x3e273d5cu, 0xbe273d5cu, 0x3d8a8bd2u}, \
    {0x3d0d42a9u, 0xbdc92352u, 0x3e1682f9u, 0xbe318a87u, 0x3e318a87u, 0xbe1682f9u, 0x3dc92352u, 0xbd0d42a9u}

// Scan ("natural") order, encoder.c:42-51.  Coefficient j of a block is the column-pass output
// of vertical frequency kScanV[j] at horizontal frequency kScanH[j] (the reference stores the
// block transposed, encoder.c:660-664, and addresses it as natural_order[j].{x,y}).
#define HYDB_SCAN_V                                                                                    \
    0, 1, 0, 0, 1, 2, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 4, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 6, 7, 6, 5, 4,  \
    3, 2, 1, 0, 1, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 3, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 5, 6, 7, 7, 6, 7
#define HYDB_SCAN_H                                                                                    \
    0, 0, 1, 2, 1, 0, 0, 1, 2, 3, 4, 3, 2, 1, 0, 0, 1, 2, 3, 4, 5, 6, 5, 4, 3, 2, 1, 0, 0, 1, 2, 3,  \
    4, 5, 6, 7, 7, 6, 5, 4, 3, 2, 1, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 3, 4, 5, 6, 7, 7, 6, 5, 6, 7, 7

// inverse of the scan order: kScanIndex[kv * 8 + kh] = j
#define HYDB_SCAN_INDEX                                                                                \
    0, 2, 3, 9, 10, 20, 21, 35, 1, 4, 8, 11, 19, 22, 34, 36, 5, 7, 12, 18, 23, 33, 37, 48,             \
    6, 13, 17, 24, 32, 38, 47, 49, 14, 16, 25, 31, 39, 46, 50, 57, 15, 26, 30, 40, 45, 51, 56, 58,     \
    27, 29, 41, 44, 52, 55, 59, 62, 28, 42, 43, 53, 54, 60, 61, 63

// hf_quant_weights (encoder.c:74-93), channel order X, Y, B
#define HYDB_HF_WEIGHTS                                                                                \
    {1969, 1969, 1969, 1962, 1969, 1962, 1655, 1885, 1885, 1655, 1397, 1610, 1704, 1610, 1397, 1178,   \
     1368, 1494, 1494, 1368, 1178, 994, 1159, 1289, 1340, 1289, 1159, 994, 839, 980, 1104, 1178,      \
     1178, 1104, 980, 839, 829, 941, 1023, 1054, 1023, 941, 829, 800, 881, 928, 928, 881,             \
     800, 755, 809, 829, 809, 755, 663, 731, 731, 663, 491, 524, 491, 349, 349, 239},                 \
    {280, 280, 280, 279, 280, 279, 245, 271, 271, 245, 214, 239, 250, 239, 214, 188,                   \
     211, 226, 226, 211, 188, 164, 185, 201, 207, 201, 185, 164, 144, 163, 178, 188,                  \
     188, 178, 163, 144, 143, 157, 168, 172, 168, 157, 143, 139, 150, 156, 156, 150,                  \
     139, 133, 140, 143, 140, 133, 125, 129, 129, 125, 116, 118, 116, 107, 107, 98},                  \
    {256, 147, 147, 85, 117, 85, 60, 78, 78, 60, 43, 56, 63, 56, 43, 43,                               \
     43, 48, 48, 43, 43, 42, 43, 43, 43, 43, 43, 42, 29, 41, 43, 43,                                  \
     43, 43, 41, 29, 29, 37, 43, 43, 43, 37, 29, 27, 33, 36, 36, 33,                                  \
     27, 24, 27, 29, 27, 24, 20, 22, 22, 20, 15, 16, 15, 10, 10, 7}

// coeff_freq_context (encoder.c:53-58)
#define HYDB_FREQ_CTX                                                                                  \
    0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 15, 16, 16, 17, 17, 18, 18, 19, 19, 20, 20, \
    21, 21, 22, 22, 23, 23, 23, 23, 24, 24, 24, 24, 25, 25, 25, 25, 26, 26, 26, 26, 27, 27, 27, 27,   \
    28, 28, 28, 28, 29, 29, 29, 29, 30, 30, 30, 30

// coeff_num_non_zero_context (encoder.c:60-66) is piecewise constant in the count of non-zeros left
HD uint32_t nnz_context(uint32_t left) {
    if (left < 2) return 0;
    if (left < 3) return 31;
    if (left < 5) return 62;
    if (left < 9) return 93;
    if (left < 13) return 123;
    if (left < 21) return 152;
    if (left < 33) return 180;
    return 206;
}

// get_non_zero_context (encoder.c:680-687)
HD uint32_t predicted_nz_context(uint32_t predicted) {
    if (predicted < 8) return predicted;
    if (predicted > 64) predicted = 64;
    return 4 + (predicted >> 1);
}

}  // namespace hydb
