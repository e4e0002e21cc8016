// hoic_step.h -- host-side launchers of the env-step pipeline (one translation unit per kernel: k_p1.cu, k_p2.cu, k_p3.cu,
// k_tail.cu; each is compiled with its own compact shared-memory layout, see dev_model.h).
#pragma once
#include <cuda_runtime.h>
struct BatchDev;
struct StepOut;
// A launch covers one interleaved part of the batch: slots part, part + nparts, part + 2 nparts, ...  hoic_step runs the
// parts on parallel streams so that the tail of one part's kernel (a few long Newton solves) overlaps the next kernel of
// another part instead of leaving SMs idle.
// group 3 (hoic_step's "tick" mode): the launch's int argument is the tick, the envs come from BatchDev.tk_list.
// group (hoic_step's "carry" mode): 0 = every env of the part; 1 = the envs whose solve of substep k - 1 finished inside the capped
// k_p3s launch (BatchDev.late_stamp[env] != k); 2 = the others, taken from the carry list and the full-tier list of substep k - 1
struct StepPart { int part, nparts, group; };
// k = substep index; -1 = the pass that recomputes the forward of the state that preceded the step
cudaError_t hoic_launch_p1(const BatchDev& B, StepPart sp, const float* actions, int k, cudaStream_t stream);
cudaError_t hoic_launch_p1w(const BatchDev& B, StepPart sp, const float* actions, int k, cudaStream_t stream);   // four envs per CTA
cudaError_t hoic_launch_p2(const BatchDev& B, StepPart sp, int k, cudaStream_t stream);
cudaError_t hoic_launch_p3(const BatchDev& B, StepPart sp, int k, cudaStream_t stream);    // full-capacity tier
cudaError_t hoic_launch_p3s(const BatchDev& B, StepPart sp, int k, cudaStream_t stream);   // small tier (most envs)
cudaError_t hoic_launch_p3sb(const BatchDev& B, StepPart sp, int k, cudaStream_t stream);  // small tier with an iteration budget (carry mode)
cudaError_t hoic_launch_p3c(const BatchDev& B, StepPart sp, int k, cudaStream_t stream);   // continuation of the solves a capped k_p3s left open
cudaError_t hoic_launch_p3fc(const BatchDev& B, StepPart sp, int k, cudaStream_t stream);  // full tier, budgeted (tick mode)
// fused kernels (k_a.cu): Newton solve of substep k + k_p1's work of substep k + 1 in one launch, per capacity tier
cudaError_t hoic_launch_a(const BatchDev& B, StepPart sp, const float* actions, int k, cudaStream_t stream);
cudaError_t hoic_launch_as(const BatchDev& B, StepPart sp, const float* actions, int k, cudaStream_t stream);
// fused k_p1 + k_p2 (k_p12.cu, opt-in experiment)
cudaError_t hoic_launch_p12(const BatchDev& B, StepPart sp, const float* actions, int k, cudaStream_t stream);
cudaError_t hoic_launch_tail(const BatchDev& B, StepPart sp, const float* actions, const StepOut& out, cudaStream_t stream);
// shared-memory attributes of the four kernels on the current device (once per device)
cudaError_t hoic_configure_p1();
cudaError_t hoic_configure_p1w();
cudaError_t hoic_configure_p2();
cudaError_t hoic_configure_p3();
cudaError_t hoic_configure_p3s();
cudaError_t hoic_configure_p3sb();
cudaError_t hoic_configure_p3c();
cudaError_t hoic_configure_p3fc();
cudaError_t hoic_configure_tail();
cudaError_t hoic_configure_p12();
cudaError_t hoic_configure_a();
cudaError_t hoic_configure_as();
