"""Mirror of the reference's ``dataloaders`` package for the encode path."""
