"""chirpgp_b200 -- B200-native (sm_100a) batched Gaussian filters and smoothers for chirp / instantaneous-
frequency estimation, behind the reference's ``chirpgp`` Python interface."""
