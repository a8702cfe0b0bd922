"""littlemcmc_b200: B200-native HMC / NUTS hot path behind littlemcmc's sampler API."""
__version__ = "0.1.0"
