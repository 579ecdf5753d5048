from .lpips import LPIPS  # noqa: F401
