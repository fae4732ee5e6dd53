"""Import-path mirror of the reference lib/ package for the entry points on the hot path."""
