// Internal interface between the C-ABI dispatcher (api.cu) and the per-network pipelines.
#pragma once
#include "../../include/jaqmc_b200.h"
#include "aug.cuh"

struct JqWfOut {
  float* logpsi;  // [W]
  float* sign;    // [W]
  float* grad;    // [W][3n]   (track only)
  float* lap;     // [W]       (track only)
  float* e_kin;   // [W]       (track only)
};

size_t jq_ferminet_ws_bytes(const jaqmc_ferminet_config* c, long long W, int track);
int jq_ferminet_forward(const jaqmc_ferminet_config* c, const jaqmc_ferminet_params* p, const jaqmc_system* sys,
                        const float* electrons, long long W, int track, void* ws, size_t ws_bytes, JqWfOut out,
                        cudaStream_t st);
