"""Command-line entry points of the path (skyeye.cli.validate)."""
