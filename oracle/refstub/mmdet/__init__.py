__version__ = "2.25.1-stub"
