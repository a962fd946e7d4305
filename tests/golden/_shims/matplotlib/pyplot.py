"""See __init__.py."""
