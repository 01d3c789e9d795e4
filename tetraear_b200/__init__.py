"""tetraear_b200 -- Blackwell-native drop-in for the TetraEar IQ demodulation hot path."""
