/* oracle/shims: intentionally empty (the reference includes it but uses nothing from it). */
