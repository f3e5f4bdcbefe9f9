"""Stand-in for the third-party UMNN package (out of scope; only the names are needed to import flowcon)."""


class NeuralIntegral:  # placeholder
    pass


class ParallelNeuralIntegral:  # placeholder
    pass
