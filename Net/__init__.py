"""Model zoo aliases with the reference's module names (Net.Densenet, Net.Resnet, ...)."""
