"""fpgan: host-side Python over the libfpg_b200.so C ABI (bindings, operators, network executors, trainers)."""
