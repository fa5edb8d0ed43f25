class ListedColormap:
    pass
