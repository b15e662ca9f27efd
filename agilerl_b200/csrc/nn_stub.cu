// temporary stubs so the ABI is complete while the network kernels land
#include "common.cuh"
extern "C" {
#define STUB(name, ...) int name(__VA_ARGS__) { b2rl::set_error(#name " not implemented yet"); return B2RL_EUNSUPPORTED; }
STUB(b2rl_net_workspace_bytes, const b2rl_net_desc*, int64_t, int, size_t*)
STUB(b2rl_noise_reset_from_normals, const b2rl_net_desc*, float*, const float*, void*)
STUB(b2rl_noise_reset_philox, const b2rl_net_desc*, float*, uint64_t, uint64_t, void*)
STUB(b2rl_noise_count, const b2rl_net_desc*, int64_t*)
STUB(b2rl_net_forward_q, const b2rl_net_desc*, const float*, const float*, int, const void*, const int64_t*, int64_t, float*, int64_t*, void*, size_t, void*)
STUB(b2rl_rainbow_loss, const b2rl_net_desc*, const b2rl_learn_cfg*, const b2rl_learn_bufs*, void*)
STUB(b2rl_rainbow_backward, const b2rl_net_desc*, const b2rl_learn_cfg*, const b2rl_learn_bufs*, void*)
STUB(b2rl_optim_step, const b2rl_net_desc*, const b2rl_learn_cfg*, const b2rl_learn_bufs*, void*)
STUB(b2rl_dqn_learn, const b2rl_net_desc*, const b2rl_learn_cfg*, const b2rl_learn_bufs*, void*)
STUB(b2rl_rainbow_learn, const b2rl_net_desc*, const b2rl_learn_cfg*, const b2rl_learn_bufs*, void*)
}
