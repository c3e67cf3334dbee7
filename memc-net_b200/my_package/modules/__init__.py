"""nn.Module wrappers of the motion-compensation ops (names fixed by the reference networks)."""
