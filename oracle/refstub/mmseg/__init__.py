"""TEST INFRASTRUCTURE ONLY — mmseg.utils.get_root_logger for reference src/model/backbone/mpvit.py:32."""
