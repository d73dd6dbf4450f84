"""Host side of the B200 engine: sector geometry, step-program compiler, C-ABI binding, device runtime."""
