/* Plain-C consumer of include/ddrl_b200.h: compiled with gcc -std=c99 and linked against libddrl_b200.so by
 * tests/test_abi.py::test_plain_c_program_links_and_calls_the_library (no Python, no torch, no CUDA headers). */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "ddrl_b200.h"

int main(void) {
    ddrl_ppo_hyper hp = {0.2f, 10.0f, 0.5f, 0.0f, 1.0f / 128.0f};
    ddrl_sgd_tail tail;
    memset(&tail, 0, sizeof tail);
    printf("abi %d\n", ddrl_abi_version());
    printf("fcnet %d %d %d\n", ddrl_fcnet_num_params(19, 2), ddrl_fcnet_num_params(43, 8), ddrl_fcnet_num_params(65, 2));
    printf("graphnet %d\n", ddrl_graphnet_num_params(4));
    printf("gae_ws %lld\n", (long long)ddrl_gae_ws_bytes(4, 16384));
    printf("train_ws %lld\n", (long long)ddrl_graphnet_train_ws_bytes(16384));
    /* argument validation happens before any CUDA call: a null pointer is DDRL_E_BADARG with a message */
    int rc = ddrl_fcnet_forward(NULL, NULL, NULL, NULL, 0.0f, 1, 1, 19, 2, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
    printf("badarg %d %s\n", rc, ddrl_last_error());
    printf("sizes %d %d\n", (int)sizeof hp, (int)sizeof tail.peer_x / (int)sizeof tail.peer_x[0]);
    /* layout of the fused-tail descriptor as a plain-C consumer sees it (the ctypes mirror in ddrl_b200/_lib.py must agree) */
    printf("tail %d %d %d %d %d\n", (int)sizeof tail, (int)offsetof(ddrl_sgd_tail, lr), (int)offsetof(ddrl_sgd_tail, peer_x),
           (int)offsetof(ddrl_sgd_tail, ll_ws), (int)offsetof(ddrl_sgd_tail, grad_acc));
    return 0;
}
