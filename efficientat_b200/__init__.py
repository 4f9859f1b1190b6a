"""EfficientAT hot path, B200-native."""
