"""Drop-in mirror of the reference's `models` package for the GAN training-step path."""
