"""Host-side mirror of the reference's interface for the hot path (Python, like the reference):
flags, data loaders + samplers, session-style model facades, evaluation drivers.  Everything
numeric is delegated to libmacr_b200.so through ``macr_b200.ops``; nothing here computes a
score, a loss or a gradient on the CPU."""
