// placeholder: FT entry points arrive in the next milestone
#include "u1_common.cuh"
extern "C" {
int u1ft_create(u1ft_handle_t*, int, const double*, size_t) { return U1_ERR_UNSUPPORTED; }
int u1ft_destroy(u1ft_handle_t) { return U1_ERR_UNSUPPORTED; }
size_t u1ft_weight_count(int) { return 0; }
size_t u1ft_workspace_bytes(u1ft_handle_t, int, int, int) { return 0; }
int u1ft_forward(u1ft_handle_t, const double*, double*, double*, int, int, void*, size_t, u1_stream_t) { return U1_ERR_UNSUPPORTED; }
int u1ft_phase(u1ft_handle_t, const double*, int, double*, int, int, void*, size_t, u1_stream_t) { return U1_ERR_UNSUPPORTED; }
int u1ft_action(u1ft_handle_t, const double*, double*, double*, double*, int, int, double, void*, size_t, u1_stream_t) { return U1_ERR_UNSUPPORTED; }
int u1ft_force(u1ft_handle_t, const double*, double*, double*, int, int, double, void*, size_t, u1_stream_t) { return U1_ERR_UNSUPPORTED; }
int u1ft_leapfrog(u1ft_handle_t, const double*, const double*, double*, double*, int, int, double, double, int, void*, size_t, u1_stream_t) { return U1_ERR_UNSUPPORTED; }
int u1ft_trajectory(u1ft_handle_t, double*, const double*, const double*, uint64_t, uint64_t, uint64_t, int, int, int, double, double, int, double*, int32_t*, double*, double*, double*, double*, void*, size_t, u1_stream_t) { return U1_ERR_UNSUPPORTED; }
}
