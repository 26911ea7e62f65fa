"""Drop-in replacements for the reference's Models/ packages (one module per stage directory)."""
