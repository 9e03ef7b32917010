"""B200-native drop-in for jax-cfd's staggered-grid FVM time step."""
