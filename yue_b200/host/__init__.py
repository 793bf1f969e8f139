"""Host-side mirror of the reference interface around the hot path: the config parsers, the
data container, the metrics and the recommender base classes, written from scratch with the same
names, argument meaning and error behaviour as the reference so the GPU recommenders run without
the reference tree (and drop into it)."""
