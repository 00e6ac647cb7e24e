"""petr_b200: B200-native PETR detection-head hot path."""
