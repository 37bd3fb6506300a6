"""gpitch_b200 -- B200-native variational-GP inner loop of gpitch."""
