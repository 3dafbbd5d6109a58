def find_introns(*a, **k):
    raise NotImplementedError("sorted_nearest.src.introns.find_introns is outside the oracle's scope")
