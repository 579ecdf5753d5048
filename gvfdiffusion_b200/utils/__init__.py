"""Host mirrors of the reference's `utils/` entry points that sit on the built path (loss_util)."""
