"""Import-path stubs for the reference's CLIP-guidance helpers (out of scope: SURVEY.md section 2, rows 10-11)."""
