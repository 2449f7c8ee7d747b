"""Drop-in mirrors of the reference's `representations` package (same names, signatures and return
conventions), computing on the GPU through libevrep.so.  Import e.g.
    from event_representation_study_b200.representations.gen1_transforms import get_item_transform
where the reference imports `representations.gen1_transforms`."""
