"""Stand-in for the slice of emukit 0.4.10 that ChampiB/CBO_with_OOP imports.  See ../README.md."""
