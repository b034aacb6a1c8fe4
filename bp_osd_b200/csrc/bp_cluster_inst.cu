// bp_cluster_inst.cu -- one translation unit per (precision, degree class) of the cluster (DSMEM) BP kernel.
// Compiled by bp_osd_b200/build.py with -DBPOSD_INST_REAL=double|float -DBPOSD_INST_DC=.. -DBPOSD_INST_DV=..
#define BPOSD_KERNELS_COMMON_ONLY
#define BPOSD_CLUSTER_INSTANTIATE
#include "bp_fast_kernel.cuh"
#include "bp_cluster_kernel.cuh"
template struct bposd::ClusterInst<BPOSD_INST_REAL, BPOSD_INST_DC, BPOSD_INST_DV>;
