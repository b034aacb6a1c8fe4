// bp_fast_inst.cu -- one translation unit per (precision, degree class) of the in-place shared-memory BP kernel.
// Compiled by bp_osd_b200/build.py with -DBPOSD_INST_REAL=double|float -DBPOSD_INST_DC=.. -DBPOSD_INST_DV=..
#define BPOSD_KERNELS_COMMON_ONLY
#define BPOSD_FAST_INSTANTIATE
#include "bp_fast_kernel.cuh"
template struct bposd::FastInst<BPOSD_INST_REAL, BPOSD_INST_DC, BPOSD_INST_DV>;
