"""Re-authored (static) autograd Functions behind the reference class names."""
