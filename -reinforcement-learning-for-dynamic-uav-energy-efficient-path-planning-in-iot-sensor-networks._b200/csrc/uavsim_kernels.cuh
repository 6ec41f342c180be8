// Umbrella header of the fused step / reset kernels (one source, two builds -- see uavsim_device.cuh):
//   uavsim_device.cuh  parameters, buffers, per-sensor device functions (RSSI, ADR, urgency, draws, observation)
//   uavsim_reset.cuh   episode statistics, new-episode device function, reset kernel
//   uavsim_step.cuh    collect action, staging pipeline, tile phases, pair kernel, step kernel
#pragma once
#include "uavsim_device.cuh"
#include "uavsim_reset.cuh"
#include "uavsim_step.cuh"
