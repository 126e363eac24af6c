"""Connectors needed by the plumbing examples (stdio, files)."""
