"""CPU oracle for the BP+OSD decode path -- test infrastructure only (see bposd_oracle.c header)."""
