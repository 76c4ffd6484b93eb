"""sim_a_splat_b200 -- B200-native renderer for sim_a_splat's per-step splat render path."""
__version__ = "0.1.0"
