"""torch_nf_b200: B200-native bijector-chain hot path behind the torch_nf API."""
