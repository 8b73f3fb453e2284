# B200-native drop-in for the reference package `architecture` (regular package, like the reference's).
