// Compile-and-link check of host/fd_host.hpp against libfd_b200.so (run by __graft_entry__.build()).
#include <cstdio>
#include "fd_host.hpp"
int main() {
    fd_config cfg;
    fd::check(fd_config_default(&cfg));
    auto a = fd::processing::generate_anchors2(16, {1.0f}, {32.0f, 16.0f}, 32, false);
    std::printf("abi %d anchors %zu first %.0f\n", fd_abi_version(), a.rows, a.v[0]);
    return (a.rows == 2 && a.v[0] == -248.0f && cfg.image_w == 640) ? 0 : 1;
}
