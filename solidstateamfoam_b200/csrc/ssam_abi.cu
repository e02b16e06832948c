// C-ABI of the B200-native constitutive + heat-source update (include/ssam_gpu.h).
// Host-side orchestration only: every number is produced by the kernels in kernels.cuh.
//
// One translation unit (the kernels are templates and static tables shared by everything below), organised as a
// unity build of the pieces under abi/, in dependency order:
//   abi/ctx.inc                context structure, error handling, allocation helpers, AoS<->SoA transfers, addressing build
//   abi/halo.inc               processor-patch halo transports: peer memory (IPC or in-process), NCCL, generalised plane exchange
//   abi/launch.inc             friction-patch sums, kernel launchers (direct, tiled, persistent), staged gradient / viscosity
//   abi/layout.inc             tile traversal order and the tile-compact cell layout (host-side integer preprocessing)
//   abi/entry_context.inc      extern "C": lifecycle, switches, profiling
//   abi/entry_mesh.inc         extern "C": mesh set-up and the integer addressing tables
//   abi/entry_staged.inc       extern "C": fields, boundary condition, staged entry points, interface, face fields, friction
//   abi/update.inc             the fused update and its multi-rank schedules
//   abi/host_pipeline.inc      host-staged entry point: simple and chunk-pipelined
//   abi/entry_misc.inc         extern "C": device-math evaluation, FP64 ceiling, communicators

#include "abi/ctx.inc"
#include "abi/halo.inc"
#include "abi/launch.inc"
#include "abi/layout.inc"
#include "abi/entry_context.inc"
#include "abi/entry_mesh.inc"
#include "abi/entry_staged.inc"
#include "abi/update.inc"
#include "abi/host_pipeline.inc"
#include "abi/entry_misc.inc"
