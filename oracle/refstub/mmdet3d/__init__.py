__version__ = "1.0.0rc4-stub"
