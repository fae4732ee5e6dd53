"""Import-path mirror of the reference agents/ package: the tabular agents, on the device."""
