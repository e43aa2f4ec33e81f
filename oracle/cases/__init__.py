"""Oracle restatements of the reference's case modules (test infrastructure)."""
