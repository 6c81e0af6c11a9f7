"""TEST INFRASTRUCTURE ONLY — the three timm symbols reference src/model/backbone/mpvit.py:21-22 imports."""
