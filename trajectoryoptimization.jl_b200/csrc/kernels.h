// kernels.h -- host-side launch wrappers (one per kernel), implemented in the .cu files next to this header.
#pragma once
#include "common.cuh"

// kernel 1: batched RK4 rollout / dual-number dynamics expansion           (rollout.cu)
cudaError_t launch_rollout(const DevProblem& P, cudaStream_t s);
cudaError_t launch_expand(const DevProblem& P, cudaStream_t s, int mode = 0);   // mode 1 / 2: only instances with acc1 == 1 / == 0
// kernel 2: cost + constraint + AL sweep                                     (sweep.cu)
cudaError_t launch_cost(const DevProblem& P, double* J, double* Jk, cudaStream_t s);
cudaError_t launch_merit(const DevProblem& P, double* J, double* viol, cudaStream_t s);
cudaError_t launch_cost_gradient(const DevProblem& P, double* grad, cudaStream_t s);
cudaError_t launch_cost_hessian(const DevProblem& P, double* hess, cudaStream_t s);
cudaError_t launch_al_expansion(const DevProblem& P, double* grad, double* hess, cudaStream_t s);
cudaError_t launch_eval_constraints(const DevProblem& P, int con, double* vals, cudaStream_t s);
cudaError_t launch_constraint_jacobians(const DevProblem& P, int con, double* jac, cudaStream_t s);
cudaError_t launch_constraint_hessians(const DevProblem& P, int con, int len, const double* lam, double* H, cudaStream_t s);
cudaError_t launch_projection(int cone, int p, int count, const double* x, double* px, int* err, cudaStream_t s);
cudaError_t launch_grad_projection(int cone, int p, int count, const double* x, double* J, int* err, cudaStream_t s);
cudaError_t launch_hess_projection(int cone, int p, int count, const double* x, const double* b, double* H, int* err, cudaStream_t s);
cudaError_t launch_al_update(const DevProblem& P, cudaStream_t s);
cudaError_t launch_reduce_merit(const DevProblem& P, const double* viol, double* out2, cudaStream_t s);
cudaError_t launch_shift_traj(const DevProblem& P, int steps, cudaStream_t s);
cudaError_t launch_gather_traj(const DevProblem& P, double* Xout, double* Uout, cudaStream_t s);
cudaError_t launch_scatter_traj(const DevProblem& P, const double* Xin, const double* Uin, cudaStream_t s);
cudaError_t launch_export_ab(const DevProblem& P, double* ABout, cudaStream_t s);
// kernel 3: Riccati backward pass                                             (riccati.cu)
cudaError_t launch_backward(const DevProblem& P, int* work_counter, cudaStream_t s);
bool riccati_small_supported(const DevProblem& P, bool any_batch);                     // riccati_small.cu: thread-per-instance pass for n <= 4, m <= 2
cudaError_t launch_backward_small(const DevProblem& P, cudaStream_t s);
// Lie-group error state + Riccati pass on a materialised expansion               (lie.cu)
cudaError_t launch_state_diff(const DevProblem& P, const double* Xbar, double* dx, cudaStream_t s);
cudaError_t launch_error_dynamics(const DevProblem& P, cudaStream_t s);
cudaError_t launch_error_expansion(const DevProblem& P, const double* gfull, const double* hfull, double* EG, double* EH, cudaStream_t s);
cudaError_t launch_backward_dense(const DevProblem& P, cudaStream_t s);
cudaError_t launch_expansion_compact(const DevProblem& P, cudaStream_t s);             // EC of every knot (P.compact)
cudaError_t launch_expand_lie(const DevProblem& P, cudaStream_t s, int mode = 0);     // [A_e B_e] straight from the dual-number RK4 step (rollout.cu)
// register-resident Riccati pass of the error-state Quadrotor + its record producers   (riccati_frag.cu)
cudaError_t launch_expansion_rec(const DevProblem& P, cudaStream_t s);                 // compact expansion -> REC[192..240) of every knot
cudaError_t launch_expansion_rec16(const DevProblem& P, cudaStream_t s, int mode);               // ... 16 lanes per knot, from the host-built term table (rollout.cu)
cudaError_t launch_trivial_columns_full(const DevProblem& P, cudaStream_t s);          // ... of the full-state [A B]
cudaError_t launch_trivial_columns(const DevProblem& P, cudaStream_t s);               // closed-form position / velocity columns of [A_e B_e], once per problem
cudaError_t launch_export_abe(const DevProblem& P, cudaStream_t s);                    // REC fragments -> ABe (col-major 12 x 16)
size_t frag_queue_ints(int B);
size_t frag_pool_doubles(int B, int N);                                                 // doubles of the speculative candidates' gain pool                                                         // ints of the kernel's work queue (allocated by the handle)
cudaError_t launch_backward_frag(const DevProblem& P, int* queue, double* pool, int* sticky_err, cudaStream_t s);
// forward pass: closed-loop rollout + merit + line search                     (forward.cu)
cudaError_t launch_forward(const DevProblem& P, cudaStream_t s);
cudaError_t launch_ladder(const DevProblem& P, cudaStream_t s);
cudaError_t launch_accept(const DevProblem& P, cudaStream_t s);
