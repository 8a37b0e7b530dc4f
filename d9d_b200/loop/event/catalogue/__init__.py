"""Catalogue of the events fired by the training and inference loops."""
